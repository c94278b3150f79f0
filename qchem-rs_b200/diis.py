"""DIIS extrapolation, restating core/src/diis.rs:18-59 step for step (host side, unchanged logic).

* samples are kept newest-first and truncated to `max_length`      (diis.rs:29-30)
* while fewer than `min_length` samples exist the newest Fock matrix is returned as is (diis.rs:32-38)
* B matrix: Frobenius dots of the error matrices, a border of +1, corner 0   (diis.rs:40-46)
* right-hand side e_n, solved through a QR factorisation                     (diis.rs:48-51)
* result sum_i c_i F_i with index 0 = newest                                 (diis.rs:52-58)
Returns None where nalgebra's `qr.solve` would (a zero diagonal entry of R).
"""
from __future__ import annotations

from collections import deque

import numpy as np


class Diis:
    def __init__(self, min_length: int, max_length: int):
        self.samples = deque()
        self.min_length = min_length
        self.max_length = max_length

    def fock(self, error: np.ndarray, fock: np.ndarray):
        self.samples.appendleft((error, fock))
        while len(self.samples) > self.max_length:
            self.samples.pop()
        n = len(self.samples)
        if n < self.min_length:
            return self.samples[0][1].copy()
        B = np.zeros((n + 1, n + 1))
        for i in range(n + 1):
            for j in range(i, n + 1):
                if i == n and j == n:
                    v = 0.0
                elif i == n or j == n:
                    v = 1.0
                else:
                    v = float(np.vdot(self.samples[i][0], self.samples[j][0]))
                B[i, j] = B[j, i] = v
        rhs = np.zeros(n + 1)
        rhs[n] = 1.0
        q, r = np.linalg.qr(B)
        if np.any(np.diag(r) == 0.0):
            return None
        c = np.linalg.solve(r, q.T @ rhs)
        out = np.zeros_like(fock)
        for i in range(n):
            out += c[i] * self.samples[i][1]
        return out
