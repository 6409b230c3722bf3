"""AutoTune of the thresholding percentile -- mirror of
/root/reference/spectralcluster/autotune.py (AutoTuneProxy :10-23, AutoTune :26-132).
Pure host control flow; every evaluation of the proxy runs the device pipeline."""

from __future__ import annotations

import enum
import math
import typing

import numpy as np

MIN_SEARCH_STEP = 1e-04


class AutoTuneProxy(enum.Enum):
  PercentileOverNME = enum.auto()       # (1 - p) / max eigengap   (Park et al. 2019)
  PercentileSqrtOverNME = enum.auto()   # sqrt(1 - p) / max eigengap (Xia et al. 2022)


class AutoTune:
  """Grid search (optionally hierarchical) of p_percentile minimising a DER proxy."""

  def __init__(self, p_percentile_min: float = 0.60, p_percentile_max: float = 0.95,
               init_search_step: float = 0.01, search_level: int = 1,
               proxy: AutoTuneProxy = AutoTuneProxy.PercentileSqrtOverNME):
    if not isinstance(proxy, AutoTuneProxy):
      raise TypeError("proxy must be an instance of AutoTuneProxy")
    self.p_percentile_min = p_percentile_min
    self.p_percentile_max = p_percentile_max
    self.search_step = init_search_step
    self.search_level = search_level
    self.proxy = proxy

  def get_percentile_range(self) -> typing.Sequence[float]:
    span = self.p_percentile_max - self.p_percentile_min
    points = int(math.ceil(span / self.search_step))
    return list(np.linspace(self.p_percentile_min, self.p_percentile_max, points))

  def update_percentile_range(self, p_percentile_min: float, p_percentile_max: float,
                              search_step: float) -> typing.Sequence[float]:
    self.p_percentile_min, self.p_percentile_max = p_percentile_min, p_percentile_max
    self.search_step = search_step
    return self.get_percentile_range()

  def tune(self, p_percentile_to_ratio: typing.Callable):
    """Returns (eigenvectors, n_clusters, best p) of the p with the smallest ratio (the first
    such p on ties).  Like the reference, narrows and stores its own range between levels."""
    grid = self.get_percentile_range()
    ratios = {}
    winner = None
    for _ in range(self.search_level):
      lowest = np.inf
      for position, p in enumerate(grid):
        if p in ratios:
          continue
        ratio, vectors, k = p_percentile_to_ratio(p)
        ratios[p] = ratio
        if ratio < lowest:
          lowest = ratio
          winner = (vectors, k, p, position)
      if len(grid) <= 1 or self.search_step < MIN_SEARCH_STEP:
        break
      grid = self.narrow(grid, winner[3])
    return winner[0], winner[1], winner[2]

  def narrow(self, grid: typing.Sequence[float], best_index: int) -> typing.Sequence[float]:
    """Next level's grid around the winner: +-max(2, len/8) points, half the step (stored)."""
    reach = max(2, len(grid) // 8)
    first = max(0, best_index - reach)
    last = min(len(grid) - 1, best_index + reach)
    return self.update_percentile_range(grid[first], grid[last], self.search_step / 2)
