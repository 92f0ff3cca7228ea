"""xsdba_b200 -- Blackwell (sm_100a) quantile-mapping hot path behind the xsdba train/adjust API."""
from .base import Grouper, parse_group  # noqa: F401
from .calendar import TimeAxis  # noqa: F401
from ._adjustment import (  # noqa: F401
    Dataset, dqm_adjust, dqm_train, eqm_train, group_quantile, group_rank, loess_trend, map_cdf, poly_trend, vecquantiles, qdm_adjust, qm_adjust,
)
from .utils import equally_spaced_nodes  # noqa: F401
from .adjustment import (  # noqa: F401
    DetrendedQuantileMapping, EmpiricalQuantileMapping, QuantileDeltaMapping, train_adjust_host,
)
from .detrending import LoessDetrend, PolyDetrend  # noqa: F401

__version__ = "0.1.0"
from .processing import escore, jitter, jitter_over_thresh, jitter_under_thresh  # noqa: F401
from .mbcn import MBCn, NpdfTransform, mbcn_adjust, mbcn_train, npdf_transform, rand_rot_matrix  # noqa: F401
from .periods import Periods, adjust_periods, stack_periods, unstack_periods  # noqa: F401
