"""Seeded synthetic speaker-turn d-vectors -- the benchmark / demo input.

Not part of the reference.  Speaker centroids on the unit sphere, labels as speaker turns of
20..199 frames (GaussianBlur assumes temporal contiguity), isotropic noise chosen so that the
intra-speaker cosine is ~`intra_cos`.  PCG64 streams are stable across NumPy versions; the test
suite pins this generator (and the oracle's identical copy) by SHA-256 of its output."""

import math

import numpy as np


def speaker_turn_dvectors(n, d, speakers, seed=0, intra_cos=0.8, turn=(20, 200),
                          return_labels=False):
  rng = np.random.default_rng(seed)
  centroids = rng.standard_normal((speakers, d))
  centroids /= np.linalg.norm(centroids, axis=1, keepdims=True)
  labels = np.empty(n, dtype=np.int64)
  at, previous = 0, -1
  while at < n:
    length = int(rng.integers(turn[0], turn[1]))
    who = int(rng.integers(0, speakers))
    if who == previous:
      who = (who + 1) % speakers
    labels[at:at + length] = who
    at += length
    previous = who
  sigma = math.sqrt((1.0 / intra_cos - 1.0) / d)
  x = centroids[labels] + sigma * rng.standard_normal((n, d))
  return (x, labels) if return_labels else x
