"""Minimal reader for R's serialization format (RDX3 / XDR), enough for ``data/pbmc3k.RData``.

The reference ships its demo dataset as a bzip2-compressed ``save()`` image whose single object
``pbmc3k`` is a named list of the dgCMatrix slots with run-length-encoded values (reference
R/get_pbmc3k_data.R:14-20 rebuilds the matrix with ``inverse.rle``). This module parses that
container without R so the fixture can be regenerated from the reference data file
(scripts/make_pbmc3k_fixture.py); see SURVEY.md App. B.2 for the layout.

Only the SEXP types that occur in that file are handled: NILSXP(254), SYMSXP(1), LISTSXP(2),
CHARSXP(9), LGLSXP(10), INTSXP(13), REALSXP(14), STRSXP(16), VECSXP(19), REFSXP(255), plus
attribute / tag flags.
"""
from __future__ import annotations

import bz2
import gzip
import struct

import numpy as np


class _Reader:
    def __init__(self, buf: bytes):
        self.b, self.o, self.refs = buf, 0, []

    def i32(self) -> int:
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def raw(self, n: int) -> bytes:
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def length(self) -> int:
        n = self.i32()
        if n == -1:  # long vector: two 32-bit halves
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def item(self):
        flags = self.i32()
        ty, has_attr, has_tag = flags & 0xFF, bool(flags & 0x200), bool(flags & 0x400)
        if ty == 254:  # NILVALUE_SXP
            return None
        if ty == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if ty == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if ty == 2:  # LISTSXP (pairlist): attr?, tag?, car, cdr
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                nxt = self.i32()
                nty = nxt & 0xFF
                if nty == 254:
                    break
                if nty != 2:
                    raise ValueError(f"unexpected pairlist tail type {nty}")
                has_attr, has_tag = bool(nxt & 0x200), bool(nxt & 0x400)
            return out
        if ty == 9:  # CHARSXP
            n = self.i32()
            return None if n == -1 else self.raw(n).decode("utf-8", "replace")
        if ty in (10, 13):  # LGLSXP / INTSXP
            n = self.length()
            v = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
        elif ty == 14:  # REALSXP
            n = self.length()
            v = np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64)
        elif ty == 16:  # STRSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif ty == 19:  # VECSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        else:
            raise ValueError(f"unsupported SEXP type {ty} at offset {self.o}")
        if has_attr:
            attrs = dict(self.item())
            if ty == 19 and "names" in attrs:
                v = {k: e for k, e in zip(attrs["names"], v)}
                v["__attrs__"] = {k: a for k, a in attrs.items() if k != "names"}
            elif isinstance(v, np.ndarray) and attrs:
                v = {"__value__": v, "__attrs__": attrs}
        return v


def read_rdata(path: str) -> dict:
    """Return {object name: value} for an ``.RData`` / ``.rda`` file."""
    with open(path, "rb") as fh:
        buf = fh.read()
    if buf[:3] == b"BZh":
        buf = bz2.decompress(buf)
    elif buf[:2] == b"\x1f\x8b":
        buf = gzip.decompress(buf)
    if buf[:5] != b"RDX3\n" and buf[:5] != b"RDX2\n":
        raise ValueError("not an RDX2/RDX3 save image")
    r = _Reader(buf)
    r.o = 5
    if r.raw(2) != b"X\n":
        raise ValueError("only XDR serialization is supported")
    version = r.i32()
    r.i32()  # writer version
    r.i32()  # min reader version
    if version == 3:
        r.raw(r.i32())  # native encoding
    top = r.item()
    return {k: v for k, v in top}


def inverse_rle(rle: dict) -> np.ndarray:
    """R's ``inverse.rle`` for an ``rle`` object parsed by :func:`read_rdata`."""
    return np.repeat(np.asarray(rle["values"]), np.asarray(rle["lengths"]))
