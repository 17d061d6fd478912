"""Drop-in for ``wmf_py/models_py/core.py``: ``ModelParams``, ``ModelState`` and
``shia_v1(forcings, grid, params, state, nsteps)`` with the reference's semantics
(core.py:24-295), computed in fp64 by the time-fused CUDA kernel ``k_shia3``."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np
import torch

from .. import _lib, ops
from .._dev import require_cuda, to_dev


@dataclass
class ModelParams:                      # core.py:24-45 (same fields, same defaults)
    dt: float
    h_coef: float
    h_exp: float
    v_coef: float
    v_exp: float
    max_capilar: float
    max_gravita: float
    max_aquifer: float
    drena: float
    retorno_gr: float
    storage_constant: float
    k_perc_cap_to_grav: float = 0.0
    k_perc_grav_to_aqu: float = 0.0
    k_baseflow: float = 0.0
    max_infil_mm_h: float = 1.0e6
    eta_priority: str = "capilar_first"
    separate_fluxes: bool = False
    save_vfluxes: bool = False
    save_storage: bool = False
    verbose: bool = False


@dataclass
class ModelState:                       # core.py:48-61
    storage_capilar: np.ndarray
    storage_gravita: np.ndarray
    storage_aquifer: np.ndarray
    q_super_last: np.ndarray | None = None
    q_base_last: np.ndarray | None = None


def shia_v1(forcings: Dict[str, np.ndarray], grid: Dict[str, np.ndarray], params: ModelParams, state: ModelState,
            nsteps: int) -> Tuple[ModelState, Dict[str, np.ndarray]]:
    """3-tank hydrological balance; inputs are not mutated; raises like the reference
    (KeyError without ``rain_mm_h``; RuntimeError when the cumulative mass error exceeds 1 %)."""
    rain = forcings.get("rain_mm_h")
    if rain is None:
        raise KeyError("forcings must contain 'rain_mm_h'")
    require_cuda()
    ny, nx = rain.shape[1:]
    ncell = ny * nx
    dt = params.dt
    rain_d = to_dev(rain, torch.float64)[:nsteps].reshape(nsteps, ncell).contiguous()
    et_d, et_scale = None, 0.0
    if "et_mm_h" in forcings:
        et_d, et_scale = to_dev(forcings["et_mm_h"], torch.float64)[:nsteps].reshape(nsteps, ncell).contiguous(), dt / 3600.0
    elif "et_mm_d" in forcings:
        et_d, et_scale = to_dev(forcings["et_mm_d"], torch.float64)[:nsteps].reshape(nsteps, ncell).contiguous(), dt / 86400.0
    cap = to_dev(state.storage_capilar, torch.float64).reshape(ncell).clone()
    grav = to_dev(state.storage_gravita, torch.float64).reshape(ncell).clone()
    aqu = to_dev(state.storage_aquifer, torch.float64).reshape(ncell).clone()
    prm = _lib.Shia3Params(dt, params.max_capilar, params.max_gravita, params.max_aquifer, params.drena,
                           params.retorno_gr, params.k_perc_cap_to_grav, params.k_perc_grav_to_aqu, params.k_baseflow,
                           params.max_infil_mm_h, et_scale, int(params.eta_priority == "proportional"), 0)
    out = ops.shia3_run(rain_d, et_d, cap, grav, aqu, prm, nsteps, separate_fluxes=params.separate_fluxes,
                        save_storage=params.save_storage)
    sums = out["sums"].cpu().numpy()                       # [nsteps+1, 6]
    # core.py:234-256 on the per-step sums {q_total, storage, P, ET, q_super, q_base}
    q_outlet = np.zeros(nsteps, dtype=float)
    mass_err_step = np.zeros(nsteps, dtype=float)
    mass_err_cum = np.zeros(nsteps, dtype=float)
    cell_area = float(grid["dx"]) * float(grid["dy"]) if ("dx" in grid and "dy" in grid) else None
    prev = sums[0, 1]
    cum_p = cum_err = 0.0
    failed = False
    for t in range(nsteps):
        q_total_sum, new_sum, P, ETsum, Qsuper, Qbase = sums[t + 1]
        q_outlet[t] = q_total_sum if cell_area is None else q_total_sum * 1.0e-3 * cell_area / dt
        deltaS = new_sum - prev
        prev = new_sum
        err = P - (deltaS + Qsuper + Qbase + ETsum)
        cum_p += P
        cum_err += err
        mass_err_step[t] = abs(err) / P if P > 0 else 0.0
        mass_err_cum[t] = abs(cum_err) / cum_p if cum_p > 0 else 0.0
        if mass_err_cum[t] > 0.01:
            failed = True
            break
        if params.verbose and ((t + 1) % max(1, nsteps // 10) == 0):
            print(f"step {t + 1}: P={P:.3f} ET={ETsum:.3f} Q={Qsuper + Qbase:.3f} dS={deltaS:.3f} "
                  f"err={mass_err_step[t]:.3%}")
    if failed:
        raise RuntimeError("cumulative mass balance error >1%")

    def host(t, shape):
        return t.cpu().numpy().reshape(shape)

    new_state = ModelState(host(cap, (ny, nx)), host(grav, (ny, nx)), host(aqu, (ny, nx)),
                           host(out["q_super_last"], (ny, nx)), host(out["q_base_last"], (ny, nx)))
    diagnostics: Dict[str, np.ndarray] = {"q_outlet": q_outlet}
    for k in ("q_super_t", "q_base_t", "q_total_t", "storage_capilar_t", "storage_gravita_t", "storage_aquifer_t"):
        if k in out:
            diagnostics[k] = host(out[k], (nsteps, ny, nx))
    diagnostics["diag"] = {"mass_error_step": mass_err_step, "mass_error_cum": mass_err_cum}
    return new_state, diagnostics
