"""Preset configurations -- mirror of /root/reference/spectralcluster/configs.py: the ICASSP 2018
"Speaker Diarization with LSTM" setup (:21-43) and the Turn-to-Diarize setup (:45-80: percentile
thresholding with binarisation, constraint propagation before refinement, AutoTune, GraphCut)."""

from . import autotune
from . import constraint
from . import laplacian
from . import refinement
from . import spectral_clusterer

AutoTune = autotune.AutoTune
ConstraintName = constraint.ConstraintName
ConstraintOptions = constraint.ConstraintOptions
LaplacianType = laplacian.LaplacianType
RefinementName = refinement.RefinementName
RefinementOptions = refinement.RefinementOptions
ThresholdType = refinement.ThresholdType
SymmetrizeType = refinement.SymmetrizeType
SpectralClusterer = spectral_clusterer.SpectralClusterer

ICASSP2018_REFINEMENT_SEQUENCE = [
    RefinementName.CropDiagonal,
    RefinementName.GaussianBlur,
    RefinementName.RowWiseThreshold,
    RefinementName.Symmetrize,
    RefinementName.Diffuse,
    RefinementName.RowWiseNormalize,
]

icassp2018_refinement_options = RefinementOptions(
    gaussian_blur_sigma=1,
    p_percentile=0.95,
    thresholding_soft_multiplier=0.01,
    thresholding_type=ThresholdType.RowMax,
    refinement_sequence=ICASSP2018_REFINEMENT_SEQUENCE)

icassp2018_clusterer = SpectralClusterer(
    min_clusters=2,
    max_clusters=7,
    autotune=None,
    laplacian_type=None,
    refinement_options=icassp2018_refinement_options,
    custom_dist="cosine")

TURNTODIARIZE_REFINEMENT_SEQUENCE = [RefinementName.RowWiseThreshold, RefinementName.Symmetrize]

turntodiarize_refinement_options = RefinementOptions(
    thresholding_soft_multiplier=0.01,
    thresholding_type=ThresholdType.Percentile,
    thresholding_with_binarization=True,
    thresholding_preserve_diagonal=True,
    symmetrize_type=SymmetrizeType.Average,
    refinement_sequence=TURNTODIARIZE_REFINEMENT_SEQUENCE)

turntodiarize_constraint_options = constraint.ConstraintOptions(
    constraint_name=constraint.ConstraintName.ConstraintPropagation,
    apply_before_refinement=True,
    constraint_propagation_alpha=0.4)

turntodiarize_auto_tune = autotune.AutoTune(
    p_percentile_min=0.40, p_percentile_max=0.95, init_search_step=0.05, search_level=1)

turntodiarize_clusterer = SpectralClusterer(
    min_clusters=2,
    max_clusters=7,
    refinement_options=turntodiarize_refinement_options,
    constraint_options=turntodiarize_constraint_options,
    autotune=turntodiarize_auto_tune,
    laplacian_type=laplacian.LaplacianType.GraphCut,
    row_wise_renorm=True,
    custom_dist="cosine")
