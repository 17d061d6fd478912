"""Object API of the hot path: ``Basin``, ``SimuBasin`` (+ ``run_shia``) and the population
evaluation behind ``Calib_NSGAII`` with the signatures of ``wmf_heavy/wmfv2.py`` (:758-877,
:1792-2106), calling the f2py-style shims ``cu`` / ``models`` with the very call expressions the
reference uses.  The reference's own classes cannot execute as shipped (methods dedented
to module level, undefined helpers, mismatched signatures); this module carries the same
signatures so the path is callable.  GIS I/O, plotting, polygons, NetCDF persistence and the DEAP
bookkeeping are out of scope."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import cu, models

Global_EPSG = 4326


def set_map_properties(ncols: int, nrows: int, xll: float, yll: float, dx: float, dy: Optional[float] = None,
                       dxp: Optional[float] = None, nodata: float = -9999.0) -> None:
    """What read_map_raster(isDEMorDIR=True) does to the module state (wmfv2.py:227-241), without GDAL."""
    cu.ncols, cu.nrows, cu.xll, cu.yll = int(ncols), int(nrows), float(xll), float(yll)
    cu.dx = float(dx)
    cu.dy = float(dx if dy is None else dy)
    cu.dxp = float(dxp if dxp is not None else 30.0)
    cu.nodata = nodata
    models.dxp = cu.dxp


def __Add_hdr_bin_2route__(path: str, storage: bool = False):
    """'<base>[.bin|.hdr]' -> ('<base>.bin', '<base>.hdr') (undefined helper of wmfv2.py:1962)."""
    base = path
    for ext in (".bin", ".hdr", ".StObin", ".StOhdr"):
        if base.endswith(ext):
            base = base[: -len(ext)]
    return (base + ".StObin", base + ".StOhdr") if storage else (base + ".bin", base + ".hdr")


class Basin:
    """wmfv2.py:758-877.  DEM and DIR are (ncols, nrows) arrays as read_map_raster returns them
    (``Mapa.T``), DIR in keypad codes; (lat, lon) are the outlet's x, y map coordinates."""

    def __init__(self, lat=0, lon=0, DEM=None, DIR=None, name='NaN', stream=None, threshold=1000, path=None,
                 useCauceMap=None):
        self.DEM = DEM
        self.DIR = DIR
        self.name = name
        self.threshold = threshold
        if path is not None:
            raise NotImplementedError("NetCDF basins (Basin(path=...)) are outside the hot path")
        if stream is not None and useCauceMap is None:           # wmfv2.py:805-811
            err = [np.sqrt((lat - i[0]) ** 2 + (lon - i[1]) ** 2) for i in stream.structure.T]
            loc = int(np.argmin(err))
            lat, lon = stream.structure[0, loc], stream.structure[1, loc]
        # === TRAZAR LA CUENCA ===  (wmfv2.py:823-828)
        self.ncells = cu.basin_find(lat, lon, DIR, cu.ncols, cu.nrows)
        self.structure = cu.basin_cut(self.ncells)
        self.DEMvec = self.Transform_Map2Basin(DEM, [cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy])
        self.DIRvec = self.Transform_Map2Basin(DIR, [cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy])

    # wmfv2.py:1151-1170
    def Transform_Map2Basin(self, Map, MapProp):
        return cu.basin_map2basin(self.structure, Map, MapProp[2], MapProp, MapProp, MapProp, cu.nodata,
                                  self.ncells, MapProp, MapProp[1])

    # wmfv2.py:1172-1199 (raster writing omitted)
    def Transform_Basin2Map(self, BasinVar, path=None, DriverFormat='GTiff', EPSG=4326):
        map_ncols, map_nrows = cu.basin_2map_find(self.structure, self.ncells)
        M, mxll, myll = cu.basin_2map(self.structure, BasinVar, map_ncols, map_nrows, self.ncells)
        return M, [map_ncols, map_nrows, mxll, myll, cu.dx, cu.dy, cu.nodata]

    # wmfv2.py:974-982
    def GetGeo_Cell_Basics(self):
        acum, longCeld, S0, Elev = cu.basin_basics(self.structure, self.DEMvec, self.DIRvec, self.ncells)
        self.CellAcum = acum
        self.CellLong = longCeld
        self.CellSlope = S0
        self.CellHeight = Elev
        self.CellCauce = np.where(self.CellAcum > self.threshold, 1, 0)

    # wmfv2.py:1076-1089
    def GetGeo_IT(self):
        acum = cu.basin_acum(self.structure, self.ncells)
        _, _, slope, _ = cu.basin_basics(self.structure, self.DEMvec, self.DIRvec, self.ncells)
        slope = np.arctan(slope)
        slope[slope == 0] = 0.0001
        return np.log((acum * cu.dxp) / np.tan(slope))


class SimuBasin(Basin):
    """wmfv2.py:1792-1922 (modelType='cells').  Sets up the ``models`` module state for shia_v1."""

    def __init__(self, lat=None, lon=None, DEM=None, DIR=None, path=None, name='NaN', stream=None, threshold=500,
                 useCauceMap=None, noData=-999, modelType='cells', SimSed=False, SimSlides=False, dt=60,
                 SaveStorage='no', SaveSpeed='no', retorno=0, SeparateFluxes='no', SeparateRain='no',
                 ShowStorage='no', SimFloods='no', controlNodos=True, storageConstant=0.001):
        if path is not None:
            raise NotImplementedError("SimuBasin(path=...) (NetCDF) is outside the hot path")
        if modelType[0] != 'c':
            raise NotImplementedError("modelType='hills' is outside the hot path (SURVEY 8(a) S1)")
        self.segunda_cuenca = False
        self.name, self.DEM, self.DIR = name, DEM, DIR
        self.modelType, self.nodata, self.threshold, self.epsg = modelType, noData, threshold, Global_EPSG
        # TRAZAR la cuenca (wmfv2.py:1862-1866)
        self.ncells = cu.basin_find(lat, lon, DIR, cu.ncols, cu.nrows)
        self.structure = cu.basin_cut(self.ncells)
        self.DEMvec = self.Transform_Map2Basin(DEM, [cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy])
        self.DIRvec = self.Transform_Map2Basin(DIR, [cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy])
        acum = cu.basin_acum(self.structure, self.ncells)                     # :1868
        models.drena = self.structure                                          # :1872
        N = self.ncells
        # Variables físicas (:1881-1893)
        models.v_coef = np.ones((4, N), np.float32)
        models.h_coef = np.ones((4, N), np.float32)
        models.v_exp = np.ones((4, N), np.float32)
        models.h_exp = np.ones((4, N), np.float32)
        models.max_capilar = np.ones((1, N), np.float32)
        models.max_gravita = np.ones((1, N), np.float32)
        models.max_aquifer = np.ones((1, N), np.float32)
        models.storage = np.zeros((5, N), np.float32)
        models.dt = dt
        models.calc_niter = 5
        models.retorno_gr = retorno
        models.verbose = 0
        models.control = np.zeros((1, N))
        self.GetGeo_Cell_Basics()                                              # :1898
        models.hill_long = self.CellLong.reshape(1, N).astype(np.float32)
        models.elem_area = np.full((1, N), cu.dxp ** 2, np.float32)
        if controlNodos:                                                       # :1899-1903
            cauce, nodos, trazado, n_nodos, n_cauce = cu.basin_stream_nod(self.structure, self.CellAcum, threshold,
                                                                           self.ncells)
            pos = np.where(nodos != 0)[0]
            models.control[0, pos] = np.arange(1, pos.size + 1)
            self.CellNodos = nodos
        self.CellAcumVec = acum
        models.control_h = np.zeros((1, N))                                    # :1906-1916
        models.sim_sediments = int(SimSed == 'si')
        models.sim_slides = int(bool(SimSlides))
        models.save_storage = int(SaveStorage == 'si')
        models.save_speed = int(SaveSpeed == 'si')
        models.separate_fluxes = int(SeparateFluxes == 'si')
        models.separate_rain = int(SeparateRain == 'si')
        models.show_storage = int(ShowStorage == 'si')
        models.sim_floods = int(SimFloods == 'si')
        models.storage_constant = storageConstant
        self.isSetGeo = False

    # wmfv2.py:1924-2040
    def run_shia(self, Calibracion, rain_path, N_intervals, start_point=1, StorageLoc=None, HspeedLoc=None,
                 path_storage=None, path_speed=None, path_conv=None, path_stra=None, path_retorno=None, kinematicN=5,
                 QsimDataFrame=True, EvpVariable='sun', EvpSerie=None, WheretoStore=None, path_vfluxes=None,
                 Dates2Save=None, FluxesDates2Save=None, path_rc=None):
        rain_pathBin, rain_pathHdr = __Add_hdr_bin_2route__(rain_path)
        N = self.ncells
        models.rain_first_point = start_point
        models.calc_niter = kinematicN
        if path_vfluxes is not None or path_speed is not None:
            raise NotImplementedError("per-step speed / vertical-flux files are outside the hot path")
        models.save_vfluxes = 0
        # storage records (wmfv2.py:1970-1982; Modelos_heavy.f90:496-500): the states after the chosen steps go to
        # records 1, 2, ... of <path_storage>.StObin.  Dates2Save: step positions (0-based) of this run, None = all
        # (set_StorageDates and the .StOhdr layout are undefined in the reference: the header here lists record, step).
        if path_storage is not None:
            models.save_storage = 1
            path_sto_bin, path_sto_hdr = __Add_hdr_bin_2route__(path_storage, storage=True)
            if Dates2Save is None:
                print('Warning: model will save states in all time steps')
                steps = np.arange(N_intervals)
            else:
                steps = np.unique(np.asarray(Dates2Save, dtype=np.int64))
                if steps.size and (steps[0] < 0 or steps[-1] >= N_intervals):
                    raise ValueError("Dates2Save holds step positions in [0, N_intervals)")
            models.guarda_cond = np.zeros(N_intervals, dtype=np.int32)
            models.guarda_cond[steps] = np.arange(1, steps.size + 1, dtype=np.int32)
            with open(path_sto_hdr, "w") as fh:
                fh.write(f"Numero de celdas: {N}\nNumero de registros: {steps.size}\nRecord, Step\n")
                for r, t in enumerate(steps):
                    fh.write(f"{r + 1}, {int(t)}\n")
        else:
            models.save_storage = 0
            path_sto_bin = 'no_save.StObin'
        if EvpSerie is not None:
            models.evpserie = np.asarray(EvpSerie, dtype=np.float32)
        elif models.evpserie is None or len(models.evpserie) < N_intervals:
            models.evpserie = np.ones(N_intervals, dtype=np.float32)
        calib = np.asarray(Calibracion, dtype=np.float32).reshape(-1)
        if calib.size == 10:                      # NSGA genes carry 10 entries (wmfv2.py:2171-2182); Fortran wants 11
            calib = np.append(calib, np.float32(1.0))
        Qsim, Qsed, Qsep, Humedad, St1, St3, Balance, Speed, Area, Alm, Qsep_byrain = models.shia_v1(
            rain_pathBin, rain_pathHdr, calib, N, np.count_nonzero(models.control) + 1,
            np.count_nonzero(models.control_h), N_intervals, StorageLoc, HspeedLoc, path_sto_bin, path_speed,
            'no_save.bin', path_conv, path_stra, "", "", path_retorno, path_rc)
        Retornos = {'Qsim': Qsim, 'Balance': Balance, 'Storage': Alm, 'Rain_Acum': models.acum_rain,
                    'Rain_hietogram': models.mean_rain}
        if np.count_nonzero(models.control_h) > 0:
            Retornos.update({'Humedad': Humedad, 'Humedad_t1': St1, 'Humedad_t2': St3})
        if QsimDataFrame:
            import pandas as pd
            ids = models.control[models.control != 0]
            Qdict = {str(int(j)): i for i, j in zip(Retornos['Qsim'][1:], ids)}
            return Retornos, pd.DataFrame(Qdict, index=np.arange(N_intervals))
        return Retornos

    def run_shia_population(self, Calibraciones: Sequence, rain_path, N_intervals, start_point=1, StorageLoc=None,
                            HspeedLoc=None, kinematicN=5, EvpSerie=None):
        """The hot part of Calib_NSGAII (wmfv2.py:2060-2106): one launch advances every parameter set of
        a population over the same basin and rain (what ``__ejec_parallel__`` + ``Pool`` were meant to do)."""
        rain_pathBin, rain_pathHdr = __Add_hdr_bin_2route__(rain_path)
        models.rain_first_point = start_point
        models.calc_niter = kinematicN
        if EvpSerie is not None:
            models.evpserie = np.asarray(EvpSerie, dtype=np.float32)
        elif models.evpserie is None or len(models.evpserie) < N_intervals:
            models.evpserie = np.ones(N_intervals, dtype=np.float32)
        cal = np.asarray([np.asarray(c, dtype=np.float32).reshape(-1) for c in Calibraciones], dtype=np.float32)
        if cal.shape[1] == 10:
            cal = np.concatenate([cal, np.ones((cal.shape[0], 1), np.float32)], axis=1)
        _, pos = models.rain_read_ascii_table(rain_pathHdr, N_intervals)
        records = models.read_rain_records(rain_pathBin, self.ncells)
        return models.shia_v1_population(cal, records, pos, self.ncells, N_intervals, StorageLoc, HspeedLoc)

    # ------------------------------------------------------------------ C1: NSGA-II calibration (wmfv2.py:2060-2106)
    def Calib_NSGAII(self, nsga_el, nodo_eval, pop_size=40, process=4, NGEN=6, MUTPB=0.5, CXPB=0.5):
        """Multi-objective calibration with the reference's signature and return value
        ``(pop, QsimPar, fitness array (2, pop_size))``.  What the reference meant to spread over ``process`` workers
        -- one ``run_shia`` per individual -- is ONE launch per generation here: every individual of the population
        advances over the same basin and rain at once (``run_shia_population``), and the two objectives of all
        hydrographs are reduced on the device (``ops.fitness_nash_peak``).  ``process`` is accepted and ignored.
        The selection operators DEAP would provide (absent from this image) are the small numpy versions below."""
        while pop_size % 4 != 0:                      # wmfv2.py:2069-2070
            pop_size += 1
        nsga_el.set_nsgaII()
        pop = nsga_el.toolbox.population(pop_size)
        QsimPar, fits = nsga_el.evaluate_population(pop, nodo_eval)
        for ind, fit in zip(pop, fits):
            ind.fitness.values = fit
        _assign_crowding(pop, nsga_el.optimiza)
        for _ in range(NGEN):
            offspring = [nsga_el.toolbox.clone(i) for i in _sel_tournament_dcd(pop, len(pop), nsga_el.optimiza)]
            for c1, c2 in zip(offspring[::2], offspring[1::2]):
                if np.random.rand() < CXPB:
                    nsga_el.toolbox.mate(c1[0], c2[0])
                    c1.fitness.values = c2.fitness.values = None
            for mut in offspring:
                if np.random.rand() < MUTPB:
                    nsga_el.toolbox.mutate(mut[0])
                    mut.fitness.values = None
            QsimPar, fits = nsga_el.evaluate_population(offspring, nodo_eval)
            for ind, fit in zip(offspring, fits):
                if not ind.fitness.valid:
                    ind.fitness.values = fit
            pop = _sel_nsga2(pop + offspring, pop_size, nsga_el.optimiza)
        return pop, QsimPar, np.array([p.fitness.values for p in pop]).T


# ---- what DEAP provides in the reference (creator / tools): individuals with a fitness, NSGA-II selection -----------
class _Fitness:
    def __init__(self):
        self.values = None
        self.crowding = 0.0
        self.rank = 0

    @property
    def valid(self):
        return self.values is not None


class Individual(list):
    """A list holding ONE calibration (the ten genes), like tools.initRepeat(creator.Individual, attr1, n=1) builds."""

    def __init__(self, *a):
        super().__init__(*a)
        self.fitness = _Fitness()


def _dominates(a, b, weights):
    wa = [w * v for w, v in zip(weights, a)]
    wb = [w * v for w, v in zip(weights, b)]
    return all(x >= y for x, y in zip(wa, wb)) and any(x > y for x, y in zip(wa, wb))


def _fronts(pop, weights):
    """Fast non-dominated sort -> list of fronts (lists of individuals), best first."""
    n = len(pop)
    S, cnt = [[] for _ in range(n)], [0] * n
    for i in range(n):
        for j in range(i + 1, n):
            if _dominates(pop[i].fitness.values, pop[j].fitness.values, weights):
                S[i].append(j); cnt[j] += 1
            elif _dominates(pop[j].fitness.values, pop[i].fitness.values, weights):
                S[j].append(i); cnt[i] += 1
    fronts, cur = [], [i for i in range(n) if cnt[i] == 0]
    while cur:
        fronts.append([pop[i] for i in cur])
        nxt = []
        for i in cur:
            for j in S[i]:
                cnt[j] -= 1
                if cnt[j] == 0:
                    nxt.append(j)
        cur = nxt
    return fronts


def _assign_crowding(pop, weights):
    for r, front in enumerate(_fronts(pop, weights)):
        for ind in front:
            ind.fitness.rank, ind.fitness.crowding = r, 0.0
        for m in range(len(weights)):
            front.sort(key=lambda i: i.fitness.values[m])
            lo, hi = front[0].fitness.values[m], front[-1].fitness.values[m]
            front[0].fitness.crowding = front[-1].fitness.crowding = float("inf")
            if np.isfinite(hi - lo) and hi > lo:
                for a, b, c in zip(front[:-2], front[1:-1], front[2:]):
                    b.fitness.crowding += (c.fitness.values[m] - a.fitness.values[m]) / (hi - lo)


def _sel_nsga2(pop, k, weights):
    out = []
    _assign_crowding(pop, weights)
    for front in _fronts(pop, weights):
        if len(out) + len(front) <= k:
            out += front
        else:
            out += sorted(front, key=lambda i: -i.fitness.crowding)[: k - len(out)]
            break
    return out


def _sel_tournament_dcd(pop, k, weights):
    """Binary tournaments on (dominance, crowding distance), as tools.selTournamentDCD runs them."""
    def duel(a, b):
        if _dominates(a.fitness.values, b.fitness.values, weights):
            return a
        if _dominates(b.fitness.values, a.fitness.values, weights):
            return b
        if a.fitness.crowding != b.fitness.crowding:
            return a if a.fitness.crowding > b.fitness.crowding else b
        return a if np.random.rand() <= 0.5 else b

    chosen = []
    while len(chosen) < k:
        p1, p2 = np.random.permutation(len(pop)), np.random.permutation(len(pop))
        for q in (p1, p2):
            for i in range(0, len(q) - 3, 4):
                chosen.append(duel(pop[q[i]], pop[q[i + 1]]))
                chosen.append(duel(pop[q[i + 2]], pop[q[i + 3]]))
    return chosen[:k]


class _Toolbox:
    pass


class nsgaii_element:
    """wmfv2.py:2108-2235, same constructor and operators.  Genes (wmfv2.py:2171-2182): evp, infil, perco, losses,
    velRun, velSub, velSup, velStream, Hu, Hg -- the first ten entries of the Fortran's calib(11); the eleventh
    (Max_aquifer scale) stays 1."""

    def __init__(self, pathLluvia, Qobs, npasos, inicio, SimuBasinElem, evp=[0, 1], infil=[1, 200], perco=[1, 40],
                 losses=[0, 1], velRun=[0.1, 1], velSub=[0.1, 1], velSup=[0.1, 1], velStream=[0.1, 1], Hu=[0.1, 1],
                 Hg=[0.1, 1], probCruce=None, probMutacion=None, rangosMutacion=None, MaxMinOptima=(1.0, -1.0),
                 CrowDist=0.5):
        self.evp_range, self.infil_range, self.perco_range, self.losses_range = evp, infil, perco, losses
        self.velRun_range, self.velSub_range, self.velSup_range, self.velStream_range = velRun, velSub, velSup, velStream
        self.hu_range, self.hg_range = Hu, Hg
        self.npasos, self.inicio, self.path_lluvia, self.Qobs, self.simelem = npasos, inicio, pathLluvia, Qobs, SimuBasinElem
        self.prob_cruce = probCruce if probCruce is not None else np.ones(10) * 0.5
        self.prob_mutacion = probMutacion if probMutacion is not None else np.ones(10) * 0.5
        self.rangos_mutacion = rangosMutacion if rangosMutacion is not None else [
            evp, infil, perco, losses, velRun, velSub, velSup, velStream, Hu, Hg]
        self.optimiza = MaxMinOptima
        self.crowdist = CrowDist

    def __crea_calibracion__(self):
        return [np.random.uniform(r[0], r[1], 1) for r in (
            self.evp_range, self.infil_range, self.perco_range, self.losses_range, self.velRun_range, self.velSub_range,
            self.velSup_range, self.velStream_range, self.hu_range, self.hg_range)]

    def __crea_ejec__(self, calibracion):
        return [calibracion, self.path_lluvia, self.npasos, self.inicio, self.simelem]

    def __evalfunc__(self, Qsim, f1=None, f2=None):
        """(Nash, peak-flow error) of one hydrograph: the device reduction on a population of one."""
        import torch
        from . import ops
        q = torch.as_tensor(np.asarray(Qsim, dtype=np.float32).reshape(1, -1)).cuda()
        o = torch.as_tensor(np.asarray(self.Qobs, dtype=np.float32).reshape(-1)[: q.shape[1]]).cuda()
        e1, e2 = ops.fitness_nash_peak(q, o)[0].tolist()
        return e1, e2

    def __cruce__(self, indi1, indi2):
        for i, p_cr in enumerate(self.prob_cruce):
            if np.random.rand() < p_cr:
                indi1[i], indi2[i] = indi2[i], indi1[i]
        return indi1, indi2

    def __mutacion__(self, indi):
        for i, p_mut in enumerate(self.prob_mutacion):
            if np.random.rand() < p_mut:
                r = self.rangos_mutacion[i]
                indi[i] = np.random.uniform(r[0], r[1], 1)
        return indi

    def set_nsgaII(self):
        import copy
        tb = _Toolbox()
        tb.attr1 = self.__crea_calibracion__
        tb.individual = lambda: Individual([tb.attr1()])
        tb.population = lambda n: [tb.individual() for _ in range(n)]
        tb.evaluate, tb.mate, tb.mutate, tb.clone = self.__evalfunc__, self.__cruce__, self.__mutacion__, copy.deepcopy
        tb.select = lambda pop, k: _sel_nsga2(pop, k, self.optimiza)
        self.toolbox = tb

    def evaluate_population(self, pop, nodo_eval):
        """Hydrographs at control row ``nodo_eval`` and (Nash, peak error) of every individual: one population launch
        and one fitness reduction (with ``models.route_channels = 1`` the routed extension runs set by set)."""
        import torch
        from . import ops
        cals = [np.concatenate([np.asarray(g, dtype=np.float32).reshape(-1) for g in ind[0]]) for ind in pop]
        T = int(self.npasos)
        if int(models.route_channels) == 1:
            Q = np.stack([self.simelem.run_shia(c, self.path_lluvia, T, start_point=self.inicio, QsimDataFrame=False)["Qsim"][nodo_eval]
                          for c in cals]).astype(np.float32)
        else:
            self.simelem.run_shia_population(cals, self.path_lluvia, T, start_point=self.inicio)
            Q = np.zeros((len(cals), T), dtype=np.float32)     # as shipped the model moves no water between cells: Q == 0
        qd = torch.from_numpy(np.ascontiguousarray(Q)).cuda()
        od = torch.from_numpy(np.asarray(self.Qobs, dtype=np.float32).reshape(-1)[:T].copy()).cuda()
        fits = ops.fitness_nash_peak(qd, od).cpu().numpy()
        return list(Q), [tuple(f) for f in fits]
