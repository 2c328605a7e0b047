"""extrack_b200 — B200-native (sm_100a CUDA) track-likelihood engine behind ExTrack's Python API.

Only the hot path of ``extrack.tracking`` is provided: ``param_fitting`` / ``cum_Proba_Cs`` /
``Proba_Cs`` / ``predict_Bs`` and the parameter builders, plus the data formats either side of it:
``readers.read_table`` (table -> length-bucketed ``all_tracks``) and ``simulate.sim_tracks``.  See DESIGN.md.
"""
from . import readers, tracking  # noqa: F401
from .tracking import (  # noqa: F401
    Proba_Cs,
    cum_Proba_Cs,
    extract_params,
    generate_params,
    get_2DSPT_params,
    get_params,
    param_fitting,
    predict_Bs,
)

__version__ = "0.1.0"
