"""CPU oracle for the SpectralClusterer.predict() hot path.

TEST INFRASTRUCTURE ONLY.  This module is a NumPy/SciPy/scikit-learn
restatement of the reference algorithm (wq2012/SpectralCluster v0.2.22).  It is
imported only by `tests/`, by `__graft_entry__.smoke()` and by the
`cpu_baseline` / `--impl reference` legs of `bench.py`, and only as the checker
or the timed CPU baseline -- never by the product package
`spectralcluster_b200`, which has no CPU fallback.

Parity pin: every function here is checked against the unmodified reference
(imported from /root/reference in the build container) by
`tests/golden/make_golden.py`, which also writes the committed fixtures under
`tests/golden/`; `tests/test_oracle_golden.py` replays them without the
reference.  The literal known-answer vectors of the reference's own unit tests
(tests/utils_test.py, refinement_test.py, laplacian_test.py,
custom_distance_kmeans_test.py, spectral_clusterer_test.py) are replayed in
`tests/test_oracle_known_answers.py`.

All arithmetic that the reference delegates to third-party code is delegated
to the same routine here (versions in this image: numpy 2.3.5, scipy 1.18.1,
scikit-learn 1.9.0 -- the reference's requirements.txt:1-3 pins none):
np.matmul, np.linalg.eig, np.percentile, scipy.ndimage.gaussian_filter,
scipy.spatial.distance.cdist, sklearn.cluster.KMeans.

Each function cites the reference lines (relative to /root/reference) it
follows.  The option bag is a plain dict so that this file shares no class
structure with the reference.
"""

from __future__ import annotations

import math

import numpy as np

EPS = 1e-10

# Option-bag defaults: refinement.py:76-100, spectral_clusterer.py:29-46.
DEFAULTS = dict(
    min_clusters=None,
    max_clusters=None,
    sequence=(),                 # names: crop, blur, threshold, symmetrize, diffuse, rownorm
    sigma=1,
    p=0.95,
    mult=0.01,
    threshold_type="rowmax",     # or "percentile"
    binarize=False,
    preserve_diagonal=False,
    symmetrize_type="max",       # or "average"
    laplacian=None,              # None | affinity | unnormalized | randomwalk | graphcut
    stop_eigenvalue=1e-2,
    row_wise_renorm=False,
    custom_dist="cosine",
    max_iter=300,
    eigengap="ratio",            # or "normalizeddiff"
    autotune=None,               # dict(p_min, p_max, step, level, proxy)
)

ICASSP2018 = ("crop", "blur", "threshold", "symmetrize", "diffuse", "rownorm")


def options(**kw):
  bag = dict(DEFAULTS)
  unknown = set(kw) - set(bag)
  if unknown:
    raise KeyError(sorted(unknown))
  bag.update(kw)
  return bag


# ---------------------------------------------------------------------------
# a2  utils.compute_affinity_matrix  (utils.py:20-41)
# ---------------------------------------------------------------------------
def affinity(x):
  unit = x / np.linalg.norm(x, axis=1)[:, None]          # utils.py:32-33
  return (np.matmul(unit, unit.T) + 1.0) / 2.0           # utils.py:35-39


# ---------------------------------------------------------------------------
# a4..a9  refinement operators  (refinement.py:136-245)
# ---------------------------------------------------------------------------
def crop_diagonal(a):
  out = np.array(a, copy=True)                           # refinement.py:147
  n = out.shape[0]
  out[np.arange(n), np.arange(n)] = 0.0                  # :148
  out[np.arange(n), np.arange(n)] = out.max(axis=1)      # :149-150
  return out


def gaussian_blur(a, sigma):
  from scipy import ndimage                              # refinement.py:162
  return ndimage.gaussian_filter(a, sigma=sigma)


def row_threshold(a, p, mult, kind="rowmax", binarize=False,
                  preserve_diagonal=False):
  out = np.array(a, copy=True)                           # refinement.py:184
  if preserve_diagonal:
    np.fill_diagonal(out, 0.0)                           # :185-186
  if kind == "rowmax":
    cut = out.max(axis=1)[:, None] * p                   # :189-191
  elif kind == "percentile":
    cut = np.percentile(out, p * 100, axis=1)[:, None]   # :194-197
  else:
    raise ValueError("Unsupported thresholding_type")    # :199
  small = out < cut
  keep = np.invert(small)
  if binarize:
    out = np.ones_like(out) * keep + out * mult * small  # :202-204
  else:
    out = out * keep + out * mult * small                # :206-207
  if preserve_diagonal:
    np.fill_diagonal(out, 1.0)                           # :208-209
  return out


def symmetrize(a, kind="max"):
  if kind == "max":
    return np.maximum(a, a.T)                            # refinement.py:221-222
  if kind == "average":
    return 0.5 * (a + a.T)                               # :223-224
  raise ValueError("Unsupported symmetrize_type.")       # :226


def diffuse(a):
  return np.matmul(a, a.T)                               # refinement.py:234


def row_normalize(a):
  out = np.array(a, copy=True)
  out /= out.max(axis=1)[:, None]                        # refinement.py:242-244
  return out


def refine(a, opt):
  """The loop of spectral_clusterer.py:131-135 over the option bag."""
  for name in opt["sequence"] or ():
    if a.ndim != 2 or a.shape[0] != a.shape[1]:          # refinement.py:52-56
      raise ValueError("affinity must be a square matrix")
    if name == "crop":
      a = crop_diagonal(a)
    elif name == "blur":
      a = gaussian_blur(a, opt["sigma"])
    elif name == "threshold":
      a = row_threshold(a, opt["p"], opt["mult"], opt["threshold_type"],
                        opt["binarize"], opt["preserve_diagonal"])
    elif name == "symmetrize":
      a = symmetrize(a, opt["symmetrize_type"])
    elif name == "diffuse":
      a = diffuse(a)
    elif name == "rownorm":
      a = row_normalize(a)
    else:
      raise ValueError("Unknown refinement operation: %s" % name)
  return a


# ---------------------------------------------------------------------------
# a10  laplacian.compute_laplacian  (laplacian.py:24-60)
# ---------------------------------------------------------------------------
def laplacian(w, kind="graphcut", eps=EPS):
  deg = np.sum(w, axis=1)                                # laplacian.py:41
  lap = np.diag(deg) - w                                 # :42
  if kind == "affinity":
    return w                                             # :45-46
  if kind == "unnormalized":
    return lap                                           # :47-48
  if kind == "randomwalk":
    return np.diag(1 / (deg + eps)).dot(lap)             # :51-53
  if kind == "graphcut":
    s = np.diag(1 / (np.sqrt(deg) + eps))                # :56
    return s.dot(lap).dot(s)                             # :57-58
  raise ValueError("Unsupported laplacian_type.")


# ---------------------------------------------------------------------------
# a11  utils.compute_sorted_eigenvectors  (utils.py:44-71)
# ---------------------------------------------------------------------------
def sorted_eig(m, descend=True):
  w, v = np.linalg.eig(m)                                # utils.py:59
  w = w.real                                             # :60
  v = v.real                                             # :61
  order = np.argsort(-w) if descend else np.argsort(w)   # :62-67
  return w[order], v[:, order]


# ---------------------------------------------------------------------------
# a12  utils.compute_number_of_clusters  (utils.py:74-130)
# ---------------------------------------------------------------------------
def number_of_clusters(w, max_clusters=None, stop_eigenvalue=1e-2,
                       eigengap="ratio", descend=True, eps=EPS):
  if eigengap not in ("ratio", "normalizeddiff"):
    raise TypeError("eigengap_type must be a EigenGapType")
  end = len(w)
  if max_clusters and max_clusters + 1 < end:            # utils.py:101-102
    end = max_clusters + 1
  best, k = 0, 0
  if descend:                                            # :116-128
    for i in range(1, end):
      if w[i - 1] < stop_eigenvalue:
        break
      if eigengap == "ratio":
        d = w[i - 1] / (w[i] + eps)
      else:
        d = (w[i - 1] - w[i]) / np.max(w)
      if d > best:
        best, k = d, i
  else:                                                  # :104-115
    for i in range(1, end - 1):
      if eigengap == "ratio":
        d = w[i + 1] / (w[i] + eps)
      else:
        d = (w[i + 1] - w[i]) / np.max(w)
      if d > best:
        best, k = d, i + 1
  return k, best


def eigenvectors_ncluster(a, opt):
  """spectral_clusterer.py:108-168 without the constraint branches."""
  a = refine(a, opt)
  lap = opt["laplacian"]
  if not lap or lap == "affinity":                       # :144-153
    w, v = sorted_eig(a)
    k, gap = number_of_clusters(w, opt["max_clusters"], opt["stop_eigenvalue"],
                                opt["eigengap"], descend=True)
  else:                                                  # :154-167
    w, v = sorted_eig(laplacian(a, lap), descend=False)
    k, gap = number_of_clusters(w, opt["max_clusters"],
                                eigengap=opt["eigengap"], descend=False)
  return w, v, k, gap


# ---------------------------------------------------------------------------
# a13  custom_distance_kmeans.run_kmeans / CustomKMeans  (:13-52, :85-141)
# ---------------------------------------------------------------------------
def seed_centroids(e, k):
  from sklearn.cluster import KMeans                     # :39-43
  km = KMeans(n_clusters=k, init="k-means++", max_iter=1, random_state=0,
              n_init="auto")
  km.fit(e)
  return km.cluster_centers_


def custom_kmeans(e, centroids, max_iter, metric="cosine", tol=0.001):
  from scipy.spatial import distance
  n = e.shape[0]
  k = centroids.shape[0]
  centroids = np.array(centroids, copy=True)
  if max_iter <= 0:
    raise ValueError("Number of iterations should be a positive number")
  if n < k:
    raise ValueError("n_samples should be >= n_clusters")
  prev = 0
  rows = np.arange(n)
  for it in range(max_iter + 1):                         # :120
    d = distance.cdist(e, centroids, metric=metric)      # :123-124
    labels = d.argmin(axis=1)                            # :125
    mean_d = np.mean(d[rows, labels])                    # :126-127
    if (mean_d <= prev and mean_d >= (1 - tol) * prev) or it == max_iter:
      break                                              # :131-133
    prev = mean_d
    for c in range(k):                                   # :136-140
      members = np.where(labels == c)[0]
      if members.any():      # NB: False for the single member [0] (quirk A.4-3)
        centroids[c] = np.mean(e[members], axis=0)
  return labels


def run_kmeans(e, k, custom_dist="cosine", max_iter=300):
  if not custom_dist:                                    # :33-36
    from sklearn.cluster import KMeans
    km = KMeans(n_clusters=k, init="k-means++", max_iter=300, random_state=0,
                n_init="auto")
    return km.predict(e)     # never fitted in the reference either: raises
  return custom_kmeans(e, seed_centroids(e, k), max_iter, custom_dist)


# ---------------------------------------------------------------------------
# a14  AutoTune  (autotune.py:58-132; closure spectral_clusterer.py:274-287)
# ---------------------------------------------------------------------------
def autotune_range(p_min, p_max, step):
  count = int(np.ceil((p_max - p_min) / step))           # autotune.py:60-62
  return list(np.linspace(p_min, p_max, count))          # :63-64


def autotune(a, opt):
  """Returns (w_best, v_best, k_best, p_best, trace) for the option bag.

  State mutation of the reference objects is not modelled: the
  oracle is called on fresh option bags.
  """
  at = opt["autotune"]
  if "threshold" not in (opt["sequence"] or ()):         # spectral_clusterer.py:268-272
    raise ValueError("AutoTune is only effective when the refinement sequence"
                     "contains RowWiseThreshold")
  p_min, p_max, step = at["p_min"], at["p_max"], at["step"]
  grid = autotune_range(p_min, p_max, step)
  seen = {}
  trace = []
  best = None
  for _ in range(at.get("level", 1)):                    # autotune.py:96
    low = np.inf
    for idx, p in enumerate(grid):
      if p in seen:
        continue
      trial = dict(opt, p=p)
      w, v, k, gap = eigenvectors_ncluster(a, trial)
      proxy = at.get("proxy", "sqrt")
      if proxy == "sqrt":                                # spectral_clusterer.py:281-282
        ratio = np.sqrt(1 - p) / gap
      elif proxy == "linear":                            # :283-284
        ratio = (1 - p) / gap
      else:
        raise ValueError("Unsupported value of AutoTuneProxy")
      seen[p] = ratio
      trace.append((p, ratio, k))
      if ratio < low:                                    # autotune.py:106-111
        low = ratio
        best = (w, v, k, p, idx)
    if not grid or len(grid) == 1 or step < 1e-4:        # :113-115
      break
    reach = max(2, len(grid) // 8)                       # :121
    lo = max(0, best[4] - reach)
    hi = min(len(grid) - 1, best[4] + reach)
    p_min, p_max = grid[lo], grid[hi]
    step = step / 2
    grid = autotune_range(p_min, p_max, step)
  return best[0], best[1], best[2], best[3], trace


# ---------------------------------------------------------------------------
# a1  SpectralClusterer.predict  (spectral_clusterer.py:201-314)
# ---------------------------------------------------------------------------
def predict(x, opt, return_details=False):
  n = x.shape[0]                                         # :222
  if not isinstance(x, np.ndarray):
    raise TypeError("embeddings must be a numpy array")
  if x.ndim != 2:
    raise ValueError("embeddings must be 2-dimensional")
  a = affinity(x)                                        # :250
  if opt["autotune"]:
    w, v, k, p_best, _ = autotune(a, opt)                # :266-289
  else:
    w, v, k, gap = eigenvectors_ncluster(a, opt)         # :292-293
  if opt["min_clusters"] is not None:
    k = max(k, opt["min_clusters"])                      # :295-296
  emb = v[:, :k]                                         # :299
  if opt["row_wise_renorm"]:
    emb = emb / np.linalg.norm(emb, axis=1, ord=2).reshape(n, 1)  # :301-305
  labels = run_kmeans(emb, k, opt["custom_dist"], opt["max_iter"])  # :309-313
  if return_details:
    return labels, dict(eigenvalues=w, n_clusters=k, spectral_embeddings=emb)
  return labels


def ordered(labels):
  """utils.enforce_ordered_labels (utils.py:133-156): first-appearance relabel."""
  labels = np.asarray(labels)
  out = labels.copy()
  table = {}
  for value in labels.tolist():
    if value not in table:
      table[value] = len(table)
  for value, new in table.items():
    out[labels == value] = new
  return out


# ---------------------------------------------------------------------------
# Synthetic speaker-turn d-vectors.  Not part of the
# reference; shared by the oracle-side tests and the benchmark so that both
# arms see the same inputs.
# ---------------------------------------------------------------------------
def synthetic_dvectors(n, d, speakers, seed=0, intra_cos=0.8,
                       turn=(20, 200), return_labels=False):
  rng = np.random.default_rng(seed)
  cent = rng.standard_normal((speakers, d))
  cent /= np.linalg.norm(cent, axis=1, keepdims=True)
  lab = np.empty(n, dtype=np.int64)
  at, prev = 0, -1
  while at < n:
    length = int(rng.integers(turn[0], turn[1]))
    s = int(rng.integers(0, speakers))
    if s == prev:
      s = (s + 1) % speakers
    lab[at:at + length] = s
    at += length
    prev = s
  noise = math.sqrt((1.0 / intra_cos - 1.0) / d)
  x = cent[lab] + noise * rng.standard_normal((n, d))
  return (x, lab) if return_labels else x
