"""Namelist-level configuration of the local analysis, mirroring the reference's
``module_config.f90`` derived types (``gts_config``, ``radar_variable_config``,
``gts_variable_config``; module_config.f90:7-34) sliced for ONE updated variable, plus the
observation-type enums of ``module_param.f90:28-57,93-97``.

``sample_namelist()`` restates the shipped ``input.nml`` (input.nml:1-170) so tests and the
benchmark run the operational settings.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Dict, List

MAX_SLOTS = 5
MAX_TYPES = 16

# module_param.f90:28-57
SOUND, SYNOP, GPSPW, METAR, SHIPS = 1, 2, 8, 10, 11
# module_param.f90:93-97
DBZ, VR, ZDR, KDP = 1, 2, 3, 4
GTS, RADAR = 0, 1

GTS_NVAR = {SYNOP: 5, SHIPS: 5, METAR: 5, SOUND: 4, GPSPW: 1}  # module_letkf_core.f90:339-418


@dataclass
class TypeConfig:
    """One observation type's settings for one variable (hclr(ivar), vclr(ivar), ...)."""
    family: int
    type: int
    use_it: bool = True
    max_lz_pts: int = 500            # module_config.f90:9,30
    hclr: float = -1.0               # km; <= 0: not used for this variable
    vclr: float = -1.0               # km; <= 0: 2-D localisation
    nvar: int = 1
    is_assim: List[bool] = field(default_factory=lambda: [True] * MAX_SLOTS)
    err_muti: List[float] = field(default_factory=lambda: [1.0] * MAX_SLOTS)  # radar: [0] = error
    err_rej: List[float] = field(default_factory=lambda: [5.0] * MAX_SLOTS)


@dataclass
class VarConfig:
    """Everything ``letkf_driver`` reads from the namelist for one ``var_update`` entry."""
    types: List[TypeConfig]
    weight_function: int = 0         # module_config.f90:58
    norain_value: float = -5.0       # module_config.f90:46
    multi_infl: float = 1.0
    use_rtpp: bool = False
    rtpp_alpha: float = 0.85
    use_rtps: bool = False
    rtps_alpha: float = 0.85
    tune_q: bool = False             # module_letkf_core.f90:252-278 (moisture / number variables)


class CTypeConfig(ctypes.Structure):
    """ctypes image of ``letkf_b200_type_config`` (include/letkf_b200.h)."""
    _fields_ = [
        ("family", ctypes.c_int32), ("type", ctypes.c_int32), ("use_it", ctypes.c_int32),
        ("max_lz_pts", ctypes.c_int32), ("hclr", ctypes.c_float), ("vclr", ctypes.c_float),
        ("nvar", ctypes.c_int32), ("is_assim", ctypes.c_int32 * MAX_SLOTS),
        ("err_muti", ctypes.c_float * MAX_SLOTS), ("err_rej", ctypes.c_float * MAX_SLOTS),
    ]


class CVarConfig(ctypes.Structure):
    """ctypes image of ``letkf_b200_var_config`` (include/letkf_b200.h)."""
    _fields_ = [
        ("ntypes", ctypes.c_int32), ("weight_function", ctypes.c_int32),
        ("norain_value", ctypes.c_float), ("multi_infl", ctypes.c_float),
        ("use_rtpp", ctypes.c_int32), ("rtpp_alpha", ctypes.c_float),
        ("use_rtps", ctypes.c_int32), ("rtps_alpha", ctypes.c_float),
        ("tune_q", ctypes.c_int32),
        ("types", CTypeConfig * MAX_TYPES),
    ]


def to_c(cfg: VarConfig) -> CVarConfig:
    if len(cfg.types) > MAX_TYPES:
        raise ValueError("too many observation types")
    c = CVarConfig()
    c.ntypes = len(cfg.types)
    c.weight_function = int(cfg.weight_function)
    c.norain_value = cfg.norain_value
    c.multi_infl = cfg.multi_infl
    c.use_rtpp = int(cfg.use_rtpp)
    c.rtpp_alpha = cfg.rtpp_alpha
    c.use_rtps = int(cfg.use_rtps)
    c.rtps_alpha = cfg.rtps_alpha
    c.tune_q = int(cfg.tune_q)
    for i, t in enumerate(cfg.types):
        ct = c.types[i]
        ct.family, ct.type, ct.use_it = t.family, t.type, int(t.use_it)
        ct.max_lz_pts, ct.hclr, ct.vclr, ct.nvar = t.max_lz_pts, t.hclr, t.vclr, t.nvar
        for s in range(MAX_SLOTS):
            ct.is_assim[s] = int(t.is_assim[s]) if s < len(t.is_assim) else 0
            ct.err_muti[s] = t.err_muti[s] if s < len(t.err_muti) else 1.0
            ct.err_rej[s] = t.err_rej[s] if s < len(t.err_rej) else 5.0
    return c


# --------------------------------------------------------------------------------------
# input.nml restated
# --------------------------------------------------------------------------------------
VAR_UPDATE = ['U', 'V', 'W', 'T', 'QVAPOR', 'QRAIN', 'QSNOW', 'QGRAUP', 'QHAIL', 'QNRAIN',
              'QNSNOW', 'QNGRAUPEL', 'QNHAIL', 'MU', 'P', 'PH']                 # input.nml:7
_TUNE_Q = {'QVAPOR', 'QRAIN', 'QSNOW', 'QGRAUP', 'QHAIL', 'QNRAIN', 'QNSNOW', 'QNGRAUPEL',
           'QNHAIL'}                                                           # core:252-278

_M = -1.0
_DBZ_H = [_M] * 5 + [8.0] * 8 + [_M] * 3            # input.nml:37
_DBZ_V = [_M] * 5 + [2.0] * 8 + [_M] * 3            # input.nml:38
_VR_H = [36., 36., 12., 24., 24.] + [_M] * 8 + [24.] * 3   # input.nml:45
_VR_V = [3.] * 5 + [_M] * 11                        # input.nml:46
_SFC_H = [50.] * 5 + [_M] * 8 + [50.] * 3           # input.nml:51,77,103
_SFC_V = [3.] * 5 + [_M] * 11                       # input.nml:52,78,104
_SND_H = [75.] * 5 + [_M] * 8 + [75.] * 3           # input.nml:129
_SND_V = [3.] * 5 + [_M] * 11                       # input.nml:130
_GTS_ASSIM = [True] * 5 + [False] * 8 + [True] * 3  # input.nml:54-72
_GPS_H = [75.] * 16                                 # input.nml:151
_MULTI_INFL = [1.6] * 5 + [1.1] * 11                # input.nml:162


def sample_namelist(var: str, weight_function: int = 0, use_gpspw: bool = False,
                    use_radar: bool = True, use_gts: bool = True) -> VarConfig:
    """VarConfig for ``var`` exactly as the shipped input.nml configures it."""
    iv = VAR_UPDATE.index(var)
    types: List[TypeConfig] = []
    if use_gts:
        for t, hc, vc in ((SOUND, _SND_H, _SND_V), (SYNOP, _SFC_H, _SFC_V), (METAR, _SFC_H, _SFC_V),
                          (SHIPS, _SFC_H, _SFC_V)):
            nv = GTS_NVAR[t]
            types.append(TypeConfig(GTS, t, True, 100, hc[iv], vc[iv], nv,
                                    [_GTS_ASSIM[iv]] * nv + [False] * (MAX_SLOTS - nv),
                                    [0.5] * MAX_SLOTS, [5.0] * MAX_SLOTS))
        if use_gpspw:  # gpspw_nml%use_it is not set in input.nml (default .false.)
            types.append(TypeConfig(GTS, GPSPW, True, 20, _GPS_H[iv], _M, 1, [True] + [False] * 4,
                                    [0.5] * MAX_SLOTS, [5.0] * MAX_SLOTS))
    if use_radar:
        types.append(TypeConfig(RADAR, DBZ, True, 300, _DBZ_H[iv], _DBZ_V[iv], 1, [True] * MAX_SLOTS,
                                [2.5] * MAX_SLOTS, [20.0] * MAX_SLOTS))   # input.nml:34-38
        types.append(TypeConfig(RADAR, VR, True, 300, _VR_H[iv], _VR_V[iv], 1, [True] * MAX_SLOTS,
                                [1.0] * MAX_SLOTS, [8.0] * MAX_SLOTS))    # input.nml:42-46
    return VarConfig(types=types, weight_function=weight_function, norain_value=-5.0,
                     multi_infl=_MULTI_INFL[iv], use_rtpp=True, rtpp_alpha=0.95, use_rtps=True,
                     rtps_alpha=0.95, tune_q=var in _TUNE_Q)


def all_sample_namelists(**kw) -> Dict[str, VarConfig]:
    return {v: sample_namelist(v, **kw) for v in VAR_UPDATE}
