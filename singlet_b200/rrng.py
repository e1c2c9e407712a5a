"""R's default random number generator (Mersenne-Twister + inversion), enough of it to reproduce the
reference's ``set.seed(s); runif(n)`` and ``.Random.seed`` reads without R.

The R callers draw ``w_init`` with ``stats::runif`` (reference R/run_nmf.R:55,
R/cross_validate_nmf.R:65, R/ard_nmf.R:87) and derive the mask seed from ``.Random.seed``
(R/cross_validate_nmf.R:84, R/ard_nmf.R:92); restated per SURVEY.md App. B.3:

* ``set.seed(s)``: scramble ``s`` 50 times with the LCG ``s = 69069*s + 1``, then fill the 625-word
  ``i_seed`` with successive LCG outputs, then set ``i_seed[0] = 624`` (``mti``).
* ``runif()``: ``genrand_int32() * 2.3283064365386963e-10`` clamped into (0, 1).
* ``.Random.seed == c(10403L, i_seed[0..624])`` as signed int32, so ``.Random.seed[[3 + r]]`` is the
  MT state word ``mt[r]`` (1-based ``r``) *of the current state*.
"""
from __future__ import annotations

import numpy as np

_I2_32M1 = 2.328306437080797e-10  # R's fixup bound: 1/(2^32 - 1)


class RRng:
    """R-compatible Mersenne-Twister. ``RRng(123).runif(5)`` equals ``set.seed(123); runif(5)``."""

    def __init__(self, seed: int = 0):
        self.set_seed(seed)

    def set_seed(self, seed: int) -> None:
        s = np.uint32(int(seed) & 0xFFFFFFFF)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = np.uint32(s * np.uint32(69069) + np.uint32(1))
            words = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = np.uint32(s * np.uint32(69069) + np.uint32(1))
                words[j] = s
        self._bitgen = np.random.MT19937()
        self._set_state(words[1:].copy(), 624)

    def _set_state(self, key: np.ndarray, pos: int) -> None:
        st = self._bitgen.state
        st["state"]["key"] = key
        st["state"]["pos"] = int(pos)
        self._bitgen.state = st

    def runif(self, n: int) -> np.ndarray:
        raw = self._bitgen.random_raw(int(n)).astype(np.float64)
        u = raw * 2.3283064365386963e-10
        # R's fixup(): keep strictly inside (0, 1)
        u[u <= 0.0] = 0.5 * _I2_32M1
        u[(1.0 - u) <= 0.0] = 1.0 - 0.5 * _I2_32M1
        return u

    def dot_random_seed(self, index_1based: int) -> int:
        """Element ``.Random.seed[[index]]`` (1-based, as R indexes it) of the CURRENT state:
        [[1]] = 10403 (RNG kind), [[2]] = mti, [[3 + r]] = mt[r] as a signed int32."""
        st = self._bitgen.state["state"]
        if index_1based == 1:
            return 10403
        if index_1based == 2:
            return int(st["pos"])
        return int(np.int32(np.uint32(st["key"][index_1based - 3])))

    def matrix_runif(self, nrow: int, ncol: int) -> np.ndarray:
        """``matrix(runif(nrow * ncol), nrow, ncol)`` (column-major fill)."""
        return self.runif(nrow * ncol).reshape((nrow, ncol), order="F")
