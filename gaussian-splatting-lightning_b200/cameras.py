"""Camera container with the fields the reference's renderers read.

Mirrors ``internal/cameras/cameras.py:14-100`` (``Camera``) and the derivations of ``Cameras.__post_init__``
(``:142-192``): fov from fx, ``world_to_camera`` stored TRANSPOSED (translation in the last row), NDC projection with
znear 0.01 / zfar 100, ``full_projection = world_to_camera @ projection``, ``camera_center = inv(w2c)[3,:3]``.
When b200gs runs inside the reference, the reference's own ``Camera`` objects are passed instead — only attribute
names matter.
"""
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor


@dataclass
class Camera:
    idx: Tensor
    R: Tensor
    T: Tensor
    fx: Tensor
    fy: Tensor
    fov_x: Tensor
    fov_y: Tensor
    cx: Tensor
    cy: Tensor
    width: Tensor
    height: Tensor
    world_to_camera: Tensor
    projection: Tensor
    full_projection: Tensor
    camera_center: Tensor
    appearance_id: Optional[Tensor] = None
    normalized_appearance_id: Optional[Tensor] = None
    time: Optional[Tensor] = None
    distortion_params: Optional[Tensor] = None
    camera_type: Optional[Tensor] = None

    def to_device(self, device):
        for name in self.__dataclass_fields__:
            v = getattr(self, name)
            if isinstance(v, torch.Tensor):
                setattr(self, name, v.to(device))
        return self

    @property
    def device(self):
        return self.R.device


def make_camera(R: Tensor, T: Tensor, fx: float, fy: float, cx: float, cy: float, width: int, height: int,
                idx: int = 0) -> Camera:
    R = R.to(torch.float32)
    T = T.to(torch.float32)
    fx_t = torch.tensor(float(fx), dtype=torch.float32)
    fy_t = torch.tensor(float(fy), dtype=torch.float32)
    w_t = torch.tensor(int(width), dtype=torch.int32)
    h_t = torch.tensor(int(height), dtype=torch.int32)
    fov_x = 2 * torch.atan((w_t / 2) / fx_t)
    fov_y = 2 * torch.atan((h_t / 2) / fy_t)
    w2c = torch.zeros(4, 4)
    w2c[:3, :3] = R
    w2c[:3, 3] = T
    w2c[3, 3] = 1.0
    w2c = w2c.transpose(0, 1).contiguous()
    znear, zfar = 0.01, 100.0
    top = torch.tan(fov_y / 2) * znear
    right = torch.tan(fov_x / 2) * znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    proj = P.transpose(0, 1).contiguous()
    full = (w2c @ proj).contiguous()
    center = torch.linalg.inv(w2c)[3, :3].contiguous()
    return Camera(
        idx=torch.tensor(idx, dtype=torch.int32), R=R, T=T, fx=fx_t, fy=fy_t, fov_x=fov_x, fov_y=fov_y,
        cx=torch.tensor(float(cx), dtype=torch.float32), cy=torch.tensor(float(cy), dtype=torch.float32),
        width=w_t, height=h_t, world_to_camera=w2c, projection=proj, full_projection=full, camera_center=center,
        appearance_id=torch.tensor(0, dtype=torch.int32), normalized_appearance_id=torch.tensor(0.0),
        time=torch.tensor(0.0), distortion_params=torch.zeros(4), camera_type=torch.tensor(0, dtype=torch.int8),
    )
