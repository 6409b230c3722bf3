"""spectralcluster_b200: the SpectralClusterer.predict() hot path of wq2012/SpectralCluster,
built from scratch for NVIDIA B200 (sm_100a).  Import surface mirrors
/root/reference/spectralcluster/__init__.py:14-43 for the names on that path."""

from . import autotune
from . import configs
from . import custom_distance_kmeans
from . import fallback_clusterer
from . import laplacian
from . import refinement
from . import spectral_clusterer
from . import utils

AutoTune = autotune.AutoTune
AutoTuneProxy = autotune.AutoTuneProxy

FallbackOptions = fallback_clusterer.FallbackOptions
SingleClusterCondition = fallback_clusterer.SingleClusterCondition
FallbackClustererType = fallback_clusterer.FallbackClustererType

LaplacianType = laplacian.LaplacianType

RefinementName = refinement.RefinementName
RefinementOptions = refinement.RefinementOptions
ThresholdType = refinement.ThresholdType
SymmetrizeType = refinement.SymmetrizeType

SpectralClusterer = spectral_clusterer.SpectralClusterer

EigenGapType = utils.EigenGapType

ICASSP2018_REFINEMENT_SEQUENCE = configs.ICASSP2018_REFINEMENT_SEQUENCE

__version__ = "0.1.0"
