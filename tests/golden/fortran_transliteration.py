"""A SECOND, independent restatement of the Fortran family of the hot path -- TEST INFRASTRUCTURE ONLY.

There is no Fortran compiler in the image, so the Fortran routines cannot be run; ``oracle/wmf_oracle.c`` restates them in
C.  To keep a transcription error from hiding in the one restatement that both the oracle and the kernels were written
against, this module transliterates the same routines a second time, straight from the Fortran text and as literally as
Python allows: 1-based indices (arrays carry an unused element 0), the loops in the order they are written, every ``real``
operation rounded to IEEE single (``numpy.float32`` scalars: no contraction, ``**`` through the platform ``powf``), whole-array
``sum`` as the sequential loop gfortran emits without -ffast-math.  It is slow (pure Python) and only ever runs on the
small cases of ``make_golden_fortran.py``; ``tests/test_oracle.py`` compares the C oracle with the vectors it produced.

Transliterated (paths relative to the reference checkout):
    coord2fil_col        wmf_heavy/cuencas_heavy.f90:355-365
    basin_find           wmf_heavy/cuencas_heavy.f90:507-560   (+ basin_cut :565-576)
    basin_basics         wmf_heavy/cuencas_heavy.f90:581-627   (basin_coordXY :639-646)
    basin_stream_nod     wmf_heavy/cuencas_heavy.f90:780-809
    shia_v1              wmf_heavy/Modelos_heavy.f90:214-254 (set-up), :359-541 (time and cell loops)
    calc_speed           wmf_heavy/Modelos_heavy.f90:907-919
Deliberate differences, all outside what well-formed inputs reach: the donor scratch ``res(2,7)`` of basin_find is
``res(2,9)`` here (the Fortran would write out of bounds with eight donors), and ``Retorned`` exists whenever
``retorno_gr > 0`` (the Fortran allocates it only for ``== 1`` and then uses it for ``> 0``).
"""
import numpy as np

f32 = np.float32


# ---------------------------------------------------------------------------------------- cuencas_heavy.f90
def coord2fil_col(x, y, xll, yll, dx, ncols, nrows):
    x, y, xll, yll, dx = (f32(v) for v in (x, y, xll, yll, dx))
    col = abs(int(np.ceil((x - xll) / dx)))
    fil = nrows - int(np.floor((y - yll) / dx))
    if col > ncols or col <= 0:
        col = -999
    if fil > nrows or fil <= 0:
        fil = -999
    return col, fil


def basin_find(coli, fili, DIR, ncols, nrows):
    """DIR(c, f), 1-based, as a (ncols + 1, nrows + 1) array.  Returns (nceldas, basin_temp (4, celTot + 1))."""
    celTot = ncols * nrows
    basin_temp = np.full((4, celTot + 1), -1, dtype=np.int64)
    basin_temp[1, celTot] = 0
    basin_temp[2, celTot] = coli
    basin_temp[3, celTot] = fili
    tenia = 1
    cont2 = 1
    col2 = coli
    row2 = fili
    res = np.zeros((3, 10), dtype=np.int64)
    while col2 > 0:
        cont3 = 0
        for kf in range(1, 4):
            for kc in range(1, 4):
                c = col2 + kc - 2
                f = row2 + kf - 2
                if 1 <= c <= ncols and 1 <= f <= nrows:
                    if DIR[c, f] == 3 * kf - kc + 1:
                        cont3 = cont3 + 1
                        res[1, cont3] = c
                        res[2, cont3] = f
        if cont3 > 0:
            for i in range(1, cont3 + 1):
                basin_temp[1, celTot - tenia + 1 - i] = cont2
                basin_temp[2, celTot - tenia + 1 - i] = res[1, i]
                basin_temp[3, celTot - tenia + 1 - i] = res[2, i]
        tenia = tenia + cont3
        paso = celTot - cont2
        col2 = basin_temp[2, paso]
        row2 = basin_temp[3, paso]
        cont2 = cont2 + 1
    return tenia, basin_temp


def basin_cut(nceldas, basin_temp, ncols, nrows):
    """basin_f(3, nceldas) as a (4, nceldas + 1) array (row 0 and column 0 unused)."""
    celTot = ncols * nrows
    basin_f = np.zeros((4, nceldas + 1), dtype=np.int64)
    basin_f[:, 1:] = basin_temp[:, celTot - nceldas + 1:]
    return basin_f


def basin_coordXY(basin_f, nceldas, xll, yll, dx, nrows):
    xll, yll, dx = f32(xll), f32(yll), f32(dx)
    X = np.zeros(nceldas + 1, dtype=f32)
    Y = np.zeros(nceldas + 1, dtype=f32)
    for i in range(1, nceldas + 1):
        X[i] = xll + dx * (f32(basin_f[2, i]) - f32(0.5))
        Y[i] = yll + dx * (f32(nrows - basin_f[3, i]) + f32(0.5))
    return X, Y


def basin_basics(basin_f, DEM, DIR, nceldas, dxP, xll, yll, dx, nrows):
    """DEM, DIR: 1-based vectors (nceldas + 1).  Returns acum, long_cel, pend, elev (1-based) and the module scalars."""
    dxP = f32(dxP)
    elev = DEM.astype(f32).copy()
    acum = np.ones(nceldas + 1, dtype=np.int64)
    long_cel = np.zeros(nceldas + 1, dtype=f32)
    pend = np.zeros(nceldas + 1, dtype=f32)
    for i in range(1, nceldas + 1):
        drenaid = nceldas - basin_f[1, i] + 1
        prueba = f32(int(DIR[i]) % 2)
        if prueba == 0.0:
            long_cel[i] = dxP
        else:
            long_cel[i] = dxP * np.sqrt(f32(2.0))
        if basin_f[1, i] != 0:
            acum[drenaid] = acum[drenaid] + acum[i]
            pend[i] = abs(f32(DEM[i]) - f32(DEM[drenaid])) / long_cel[i]
        else:
            pend[i] = pend[i - 1]
        if pend[i] == 0:
            pend[i] = f32(0.001)
    area = f32(nceldas) * dxP ** 2 / f32(1e6)
    s = f32(0.0)
    for i in range(1, nceldas + 1):
        s = s + pend[i]
    pend_media = s / f32(nceldas)
    s = f32(0.0)
    for i in range(1, nceldas + 1):
        s = s + f32(DEM[i])
    elevacion = s / f32(nceldas)
    X, Y = basin_coordXY(basin_f, nceldas, xll, yll, dx, nrows)
    Xs = np.concatenate([[f32(0)], np.sort(X[1:])]).astype(f32)       # call QsortC(X)
    Ys = np.concatenate([[f32(0)], np.sort(Y[1:])]).astype(f32)
    if nceldas % 2 == 0:
        centroX, centroY = Xs[nceldas // 2], Ys[nceldas // 2]
    else:
        centroX, centroY = Xs[(nceldas + 1) // 2], Ys[(nceldas + 1) // 2]
    return acum, long_cel, pend, elev, (area, pend_media, elevacion, centroX, centroY)


def basin_stream_nod(basin_f, acum, nceldas, umbral):
    cauce = np.zeros(nceldas + 1, dtype=np.int64)
    for i in range(1, nceldas + 1):
        if acum[i] >= umbral:
            cauce[i] = 1
    n_cauce = int((cauce[1:] == 1).sum())
    nodos = np.zeros(nceldas + 1, dtype=np.int64)
    for i in range(1, nceldas + 1):
        drenaid = nceldas - basin_f[1, i] + 1
        if basin_f[1, i] != 0 and cauce[i] == 1:
            nodos[drenaid] = nodos[drenaid] + 1
    for i in range(1, nceldas + 1):
        if nodos[i] == 0 and cauce[i] == 1:
            nodos[i] = 3
    nodos[nceldas] = 2
    for i in range(1, nceldas + 1):
        if nodos[i] < 2:
            nodos[i] = 0
    n_nodos = int((nodos[1:] > 0).sum())
    trazado = np.zeros(nceldas + 1, dtype=np.int64)
    trazado[nceldas] = 1
    for i in range(1, nceldas + 1):
        drenaid = nceldas - basin_f[1, i] + 1
        if basin_f[1, i] != 0 and nodos[drenaid] != 0 and cauce[i] == 1:
            trazado[i] = 1
    return cauce, nodos, trazado, n_nodos, n_cauce


# ---------------------------------------------------------------------------------------- Modelos_heavy.f90
def _sum_seq(a):
    """sum(A) of a Fortran array: one running single-precision sum in array-element (column-major) order."""
    s = f32(0.0)
    for v in np.asarray(a, dtype=f32).ravel(order="F"):
        s = s + v
    return s


def calc_speed(sm, coef, expo, elem_long, speed, dt, calc_niter):
    """Modelos_heavy.f90:907-919 -> (speed, area)."""
    area = f32(0.0)
    for _ in range(calc_niter):
        area = sm / (elem_long + speed * dt)
        new_speed = coef * (area ** expo)
        speed = (f32(2) * new_speed + speed) / f32(3)
    return speed, area


def shia_v1(N_cel, N_reg, calib, v_coef, h_coef, h_exp, Max_capilar, Max_gravita, Max_aquifer, hill_long, elem_area,
            rain_records, posEvento, EvpSerie, dt, speed_type, retorno_gr, retorno_aq, calc_niter, StoIn=None, HspeedIn=None,
            Storage=None, storage_constant=0.001, show_storage=0, show_mean_speed=0, show_mean_retorno=0, save_vfluxes=0,
            save_rc=0, save_speed=0, save_retorno=0, step_records=None):
    """The time and cell loops as shipped (no routing: tank 5, drena and Q are never touched).

    Per-cell tables are (k, N_cel) arrays indexed [i - 1, celda - 1] where the Fortran says (i, celda); rain record r is
    rain_records[r - 1].  Returns a dict with StoOut (5, N), hspeed (4, N), balance, Mean_Rain, Acum_rain and the optional
    series the flags ask for."""
    calib = np.asarray(calib, dtype=f32)
    dt = f32(dt)
    A = lambda a: np.asarray(a, dtype=f32)
    v_coef, h_coef, h_exp = A(v_coef), A(h_coef), A(h_exp)
    hill_long, elem_area = A(hill_long).reshape(-1), A(elem_area).reshape(-1)
    Mean_Rain = np.zeros(N_reg, dtype=f32)
    Acum_rain = np.zeros(N_cel, dtype=f32)
    m3_mmHill = (elem_area / f32(1000.0)).astype(f32)
    if StoIn is not None and A(StoIn)[0, 0] > 0:
        StoOut = A(StoIn).copy()
    else:
        base = np.zeros((5, N_cel), dtype=f32) if Storage is None else A(Storage)
        StoOut = (base + f32(storage_constant)).astype(f32)
    entradas = f32(0.0)
    salidas = f32(0.0)
    balance = np.zeros(N_reg, dtype=f32)
    vspeed = np.zeros((4, N_cel), dtype=f32)
    hspeed = np.zeros((4, N_cel), dtype=f32)
    st4 = list(speed_type) + [1] * (4 - len(speed_type))
    if HspeedIn is not None and A(HspeedIn)[0, 0] > 0:
        hspeed = A(HspeedIn).copy()
    for i in range(1, 5):                                 # (vspeed is set regardless: see the oracle's header)
        vspeed[i - 1, :] = v_coef[i - 1, :] * calib[i - 1] * dt
        if not (HspeedIn is not None and A(HspeedIn)[0, 0] > 0):
            if st4[i - 1] == 1:
                hspeed[i - 1, :] = h_coef[i - 1, :] * calib[i + 4 - 1]
            else:
                hspeed[i - 1, :] = f32(0.0)
    H = np.zeros((3, N_cel), dtype=f32)
    H[0, :] = A(Max_capilar).reshape(-1) * calib[9 - 1]
    H[1, :] = A(Max_gravita).reshape(-1) * calib[10 - 1]
    H[2, :] = A(Max_aquifer).reshape(-1) * calib[11 - 1]
    Retorned = np.zeros(N_cel, dtype=f32)
    mean_retorno = np.zeros(N_reg, dtype=f32)
    mean_storage = np.zeros((5, N_reg), dtype=f32)
    mean_speed = np.zeros((4, N_reg), dtype=f32)
    vfluxes = np.zeros((4, N_cel), dtype=f32)
    mean_vfluxes = np.zeros((4, N_reg), dtype=f32)
    rc_coef = np.zeros((2, N_cel), dtype=f32)
    vflux = np.zeros(5, dtype=f32)
    hflux = np.zeros(5, dtype=f32)
    for tiempo in range(1, N_reg + 1):
        StoAtras = _sum_seq(StoOut)
        if posEvento[tiempo - 1] == 1:
            Rain = np.zeros(N_cel, dtype=f32)
        else:
            RainInt = rain_records[posEvento[tiempo - 1] - 1]
            Rain = (RainInt.astype(f32) / f32(1000.0)).astype(f32)
        rain_sum = f32(0.0)
        Acum_rain[:] = Acum_rain + Rain
        for celda in range(1, N_cel + 1):
            c = celda - 1
            entradas = entradas + Rain[c]
            rain_sum = rain_sum + Rain[c]
            vflux[1] = max(f32(0.0), Rain[c] - H[0, c] + StoOut[0, c])
            StoOut[0, c] = StoOut[0, c] + Rain[c] - vflux[1]
            Evp_loss = min(f32(EvpSerie[tiempo - 1]) * vspeed[0, c] * (StoOut[0, c] / H[0, c]) ** f32(0.6), StoOut[0, c])
            StoOut[0, c] = StoOut[0, c] - Evp_loss
            for i in range(1, 4):
                vflux[i + 1] = min(vflux[i], vspeed[i + 1 - 1, c])
                StoOut[i + 1 - 1, c] = StoOut[i + 1 - 1, c] + vflux[i] - vflux[i + 1]
            if retorno_aq > 0:
                Ret_aq = max(f32(0.0), StoOut[3, c] - H[2, c])
                StoOut[2, c] = StoOut[2, c] + Ret_aq
                StoOut[3, c] = StoOut[3, c] - Ret_aq
            if retorno_gr > 0:
                Ret = max(f32(0.0), StoOut[2, c] - H[1, c])
                StoOut[1, c] = StoOut[1, c] + Ret
                StoOut[2, c] = StoOut[2, c] - Ret
                Retorned[c] = Retorned[c] + Ret
                vflux[2] = vflux[2] + Ret
                vflux[3] = vflux[3] - Ret
            if save_vfluxes == 1:
                for i in range(1, 5):
                    vfluxes[i - 1, c] = vflux[i]
            if save_rc == 1:
                rc_coef[0, c] = rc_coef[0, c] + vflux[2] - vflux[3]
                rc_coef[1, c] = rc_coef[1, c] + Rain[c]
            salidas = salidas + vflux[4] + Evp_loss
            for i in range(1, 4):
                if st4[i - 1] == 1:
                    hflux[i] = (f32(1) - hill_long[c] / (hspeed[i - 1, c] * dt + hill_long[c])) * StoOut[i + 1 - 1, c]
                elif st4[i - 1] == 2:
                    sp, section_area = calc_speed(StoOut[i + 1 - 1, c] * m3_mmHill[c], h_coef[i - 1, c] * calib[i + 4 - 1],
                                                  h_exp[i - 1, c], hill_long[c], hspeed[i - 1, c], dt, calc_niter)
                    hspeed[i - 1, c] = sp
                    hflux[i] = min(section_area * hspeed[i - 1, c] * dt / m3_mmHill[c], StoOut[i + 1 - 1, c])
                StoOut[i + 1 - 1, c] = StoOut[i + 1 - 1, c] - hflux[i]
        Mean_Rain[tiempo - 1] = rain_sum / f32(N_cel)
        if save_vfluxes == 1:
            for i in range(4):
                mean_vfluxes[i, tiempo - 1] = _sum_seq(vfluxes[i]) / f32(N_cel)
            if step_records is not None:                 # :504-506 (the caller picks the steps guarda_vfluxes marks)
                step_records.setdefault("vfluxes", []).append(vfluxes.copy())
        if save_speed == 1 and step_records is not None:         # :509-511 write_float_basin(ruta_speed, hspeed, tiempo, ...)
            step_records.setdefault("speed", []).append(hspeed.copy())
        if save_retorno == 1 and step_records is not None:       # :513-515 write_float_basin(ruta_retorno, Retorned, tiempo, ...)
            step_records.setdefault("retorno", []).append(Retorned.copy())
        if show_storage == 1:
            for i in range(5):
                mean_storage[i, tiempo - 1] = _sum_seq(StoOut[i]) / f32(N_cel)
        if show_mean_speed == 1:
            for i in range(4):
                mean_speed[i, tiempo - 1] = _sum_seq(hspeed[i]) / f32(N_cel)
        if show_mean_retorno == 1:
            mean_retorno[tiempo - 1] = _sum_seq(Retorned) / f32(N_cel)
        if show_mean_retorno == 1 or save_retorno == 1:          # :529-531
            Retorned[:] = f32(0.0)
        balance[tiempo - 1] = _sum_seq(StoOut) - StoAtras - entradas + salidas
        entradas = f32(0.0)
        salidas = f32(0.0)
    if save_rc == 1:
        over = rc_coef[0] > rc_coef[1]
        rc_coef[0, over] = rc_coef[1, over]
    return {"StoOut": StoOut, "hspeed": hspeed, "balance": balance, "Mean_Rain": Mean_Rain, "Acum_rain": Acum_rain,
            "mean_storage": mean_storage, "mean_speed": mean_speed, "mean_retorno": mean_retorno, "vfluxes": vfluxes,
            "mean_vfluxes": mean_vfluxes, "rc_coef": rc_coef, "Retorned": Retorned}
