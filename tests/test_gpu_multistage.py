"""GPU: MultiStageClusterer.streaming_predict (multi_stage_clusterer.py:125-180) with the device
SpectralClusterer as its main stage, against the reference's labels at ten points of a 130-step
stream (tests/golden/make_golden_callers.py), for every deflicker mode."""

import os

import numpy as np
import pytest

import spectralcluster_b200 as scb
from spectralcluster_b200 import multi_stage_clusterer as ms

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "callers")
ordered = scb.utils.enforce_ordered_labels


@pytest.mark.parametrize("mode", ["NoDeflicker", "OrderBased", "Hungarian"])
def test_streaming_matches_reference(mode):
  z = np.load(os.path.join(GOLDEN, "multistage.npz"))
  x = z["x"]
  main = scb.SpectralClusterer(
      min_clusters=1, max_clusters=5,
      refinement_options=scb.RefinementOptions(
          gaussian_blur_sigma=0, p_percentile=0.9,
          refinement_sequence=list(scb.ICASSP2018_REFINEMENT_SEQUENCE)))
  clusterer = ms.MultiStageClusterer(main, fallback_threshold=0.5, L=8, U1=30, U2=60,
                                     deflicker=getattr(ms.Deflicker, mode))
  for i in range(len(x)):
    labels = clusterer.streaming_predict(x[i])
    if i + 1 in z["checkpoints"]:
      want = z["%s_%d" % (mode, i + 1)]
      if mode == "NoDeflicker":       # raw k-means ids: compare as partitions
        np.testing.assert_array_equal(ordered(np.asarray(labels)), ordered(want), err_msg=str(i + 1))
      else:
        np.testing.assert_array_equal(np.asarray(labels), want, err_msg=str(i + 1))
