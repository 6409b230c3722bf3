import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
  # The C-ABI library is a build artefact (git-ignored): make sure it exists and is current
  # before any test imports it (incremental: a no-op when the sources are unchanged).
  from spectralcluster_b200 import build
  build.build()


def golden_names():
  return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def load_golden(name):
  """Fixture written by tests/golden/make_golden.py from the unmodified reference."""
  from oracle import spectral_oracle as orc
  z = np.load(os.path.join(GOLDEN, name + ".npz"))
  case = {k: z[k] for k in z.files}
  case["options"] = eval(str(z["options"]), {"__builtins__": {}}, {})
  if "embeddings" not in case:
    n, d, speakers, seed = (int(v) for v in z["synthetic_args"])
    x = orc.synthetic_dvectors(n, d, speakers, seed=seed)
    import hashlib
    digest = hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
    assert digest == str(z["embeddings_sha256"]), "synthetic generator drifted from the fixture"
    case["embeddings"] = x
  return case


@pytest.fixture(scope="session")
def engine():
  from spectralcluster_b200 import device
  return device.Engine.get()
