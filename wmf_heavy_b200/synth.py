"""Synthetic inputs of SURVEY 8(d), generated on the device (input generation only: torch tensor ops are plumbing here,
none of this is on the measured path).

    terrain_dir   secondary stress set (a): D8 steepest descent of a tilted plane + smooth multi-scale noise, all eight
                  directions present in nearly every tile, pits where no neighbour is lower (code -1)
    nodata_mask   a random mask with a given fraction of inactive cells
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_DR = (0, 1, 1, 1, 0, -1, -1, -1)      # internal code k -> (drow, dcol), wmf_py/cu_py/basics.py:37-46
_DC = (1, 1, 0, -1, -1, -1, 0, 1)


def terrain(ny: int, nx: int, seed: int = 5, device=None, tilt=(0.7, 0.3), octaves=((8, 8.0), (48, 48.0), (256, 200.0))
            ) -> torch.Tensor:
    """z float32 [ny][nx]: a plane falling towards +row / +col plus bilinear-upsampled random grids (cell, amplitude)."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=device).manual_seed(int(seed))
    r = torch.arange(ny, device=device, dtype=torch.float32).view(ny, 1)
    c = torch.arange(nx, device=device, dtype=torch.float32).view(1, nx)
    z = -(tilt[0] * r + tilt[1] * c)
    for cell, amp in octaves:
        gy, gx = ny // cell + 3, nx // cell + 3
        coarse = torch.randn((1, 1, gy, gx), device=device, generator=g)
        up = F.interpolate(coarse, size=(gy * cell, gx * cell), mode="bicubic", align_corners=False)[0, 0]
        z = z + amp * up[cell:cell + ny, cell:cell + nx]
        del up
    return z


def steepest_descent(z: torch.Tensor) -> torch.Tensor:
    """int8 [ny][nx] internal D8 codes of the steepest downhill neighbour; -1 where none is lower (pit) ."""
    ny, nx = z.shape
    zp = F.pad(z[None, None], (1, 1, 1, 1), value=float("inf"))[0, 0]
    best = torch.zeros_like(z)
    code = torch.full((ny, nx), -1, dtype=torch.int8, device=z.device)
    for k in range(8):
        nb = zp[1 + _DR[k]:1 + _DR[k] + ny, 1 + _DC[k]:1 + _DC[k] + nx]
        slope = (z - nb) * (1.0 if _DR[k] == 0 or _DC[k] == 0 else 0.70710678)
        upd = slope > best                      # (a neighbour outside the raster is +inf: slope -inf, never chosen)
        best = torch.where(upd, slope, best)
        code = torch.where(upd, torch.full_like(code, k), code)
        del nb, slope, upd
    return code


def terrain_dir(ny: int, nx: int, seed: int = 5, device=None) -> torch.Tensor:
    return steepest_descent(terrain(ny, nx, seed, device))


def nodata_mask(ny: int, nx: int, frac: float, seed: int = 9, device=None) -> torch.Tensor:
    """uint8 [ny][nx]: 1 = active; a fraction `frac` of the cells, chosen at random, is 0."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=device).manual_seed(int(seed))
    return (torch.rand((ny, nx), device=device, generator=g) >= frac).to(torch.uint8)
