"""CPU: the C-ABI library builds, loads and exports every symbol include/*.h declares, with
the argument counts the ctypes table binds; host-side mirrors behave like the reference's
host-side logic.  No compute calls (no GPU here)."""

import os
import re

import numpy as np
import pytest

import spectralcluster_b200 as scb
from spectralcluster_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "spectralcluster_b200.h")


def declared_functions():
  text = open(HEADER).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  out = {}
  for m in re.finditer(r"\b(?:int|long long|const char\*)\s+(sc_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
    args = m.group(2).strip()
    out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
  return out


def test_library_exports_every_declared_symbol():
  lib = nat.load()          # raises if the .so is missing: there is no fallback
  decl = declared_functions()
  assert len(decl) >= 20
  for name, argc in decl.items():
    assert hasattr(lib, name), "missing export " + name
    assert name in nat.PROTOTYPES, "ctypes table lacks " + name
    assert len(nat.PROTOTYPES[name]) == argc, "arity mismatch for " + name
  assert set(nat.PROTOTYPES) == set(decl)
  assert lib.sc_abi_version() == nat.ABI_VERSION == 2


def test_no_cuda_device_is_an_error_not_a_fallback():
  import torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  with pytest.raises(RuntimeError, match="no CPU fallback"):
    scb.configs.icassp2018_clusterer.predict(np.random.rand(10, 4))


def test_api_surface_matches_reference_names():
  for name in ("AutoTune", "AutoTuneProxy", "FallbackOptions", "SingleClusterCondition",
               "FallbackClustererType", "LaplacianType", "RefinementName", "RefinementOptions",
               "ThresholdType", "SymmetrizeType", "SpectralClusterer", "EigenGapType",
               "ICASSP2018_REFINEMENT_SEQUENCE"):
    assert hasattr(scb, name)
  c = scb.SpectralClusterer()
  for attr, default in (("min_clusters", None), ("max_clusters", None), ("autotune", None),
                        ("laplacian_type", None), ("stop_eigenvalue", 1e-2),
                        ("row_wise_renorm", False), ("custom_dist", "cosine"), ("max_iter", 300),
                        ("constraint_options", None), ("max_spectral_size", None)):
    assert getattr(c, attr) == default
  assert c.eigengap_type == scb.EigenGapType.Ratio
  ro = scb.RefinementOptions()
  assert (ro.gaussian_blur_sigma, ro.p_percentile, ro.thresholding_soft_multiplier) == (1, 0.95, 0.01)
  assert ro.thresholding_type == scb.ThresholdType.RowMax
  assert ro.symmetrize_type == scb.SymmetrizeType.Max and ro.refinement_sequence is None
  assert [m.name for m in scb.RefinementName] == [
      "CropDiagonal", "GaussianBlur", "RowWiseThreshold", "Symmetrize", "Diffuse",
      "RowWiseNormalize"]
  cfg = scb.configs.icassp2018_clusterer
  assert (cfg.min_clusters, cfg.max_clusters, cfg.laplacian_type) == (2, 7, None)


def test_host_logic_matches_oracle():
  from oracle import spectral_oracle as orc
  rng = np.random.default_rng(5)
  w = np.sort(rng.random(30))[::-1]
  for mc in (None, 3, 7, 100):
    for gap, kind in (("ratio", scb.EigenGapType.Ratio),
                      ("normalizeddiff", scb.EigenGapType.NormalizedDiff)):
      for descend in (True, False):
        ww = w if descend else w[::-1]
        assert (scb.utils.compute_number_of_clusters(ww, mc, 0.05, kind, descend) ==
                orc.number_of_clusters(ww, mc, 0.05, gap, descend))
  lab = rng.integers(0, 9, 200)
  np.testing.assert_array_equal(scb.utils.enforce_ordered_labels(lab), orc.ordered(lab))
  for lo, hi, step in ((0.6, 0.95, 0.05), (0.6, 0.95, 0.045), (0.4, 0.9, 0.1)):
    assert scb.AutoTune(lo, hi, step).get_percentile_range() == orc.autotune_range(lo, hi, step)
  with pytest.raises(TypeError):
    scb.utils.compute_number_of_clusters(w, eigengap_type="ratio")
  with pytest.raises(TypeError):
    scb.AutoTune(proxy="x")
  with pytest.raises(ValueError):
    scb.RefinementOptions().get_refinement_operator("nope")


def test_autotune_search_matches_oracle_control_flow():
  """AutoTune.tune vs the oracle's restatement on a synthetic ratio curve (two levels)."""
  def curve(p):
    return (p - 0.77) ** 2 + 0.1, None, 3
  at = scb.AutoTune(0.4, 0.95, 0.05, search_level=2)
  _, _, best = at.tune(curve)
  grid = np.linspace(0.4, 0.95, int(np.ceil(0.55 / 0.05)))
  first = grid[np.argmin((grid - 0.77) ** 2)]
  assert abs(best - 0.77) <= abs(first - 0.77) + 1e-12
  assert at.search_step == 0.0125         # halved after every level, as the reference (A.4-2)


def test_predict_input_validation_order():
  c = scb.SpectralClusterer()
  with pytest.raises(AttributeError):
    c.predict([[1.0, 2.0]])               # .shape is touched before the type check (A.4-1)
  with pytest.raises(ValueError):
    c.predict(np.zeros(3))


def test_package_generator_equals_oracle_generator():
  """bench.py feeds the GPU arm from spectralcluster_b200.synthetic and the CPU arm from the
  oracle's copy: they must be the same inputs."""
  from oracle import spectral_oracle as orc
  from spectralcluster_b200 import synthetic
  for n, d, k, seed in ((777, 32, 3, 0), (2048, 256, 6, 5)):
    a, la = synthetic.speaker_turn_dvectors(n, d, k, seed=seed, return_labels=True)
    b, lb = orc.synthetic_dvectors(n, d, k, seed=seed, return_labels=True)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(la, lb)


def test_product_never_imports_the_oracle():
  pkg = os.path.join(ROOT, "spectralcluster_b200")
  for name in os.listdir(pkg):
    if name.endswith(".py"):
      text = open(os.path.join(pkg, name)).read()
      assert "oracle" not in text.replace("the oracle's identical copy", ""), name
