"""Minimal reader for R's gzip'd RDX2/XDR serialisation (``save()`` / ``save.image()`` files).

The reference ships its datasets and stored results as ``.RData`` files
(``inst/reproduce/*/data/*.RData``).  R is not available in
this environment, so this ~100-line pure-Python parser reads the subset of the
format those files use: pairlists, lists (VECSXP), real/integer/logical/string
vectors, attributes (dim, names) and symbols/references.  Closures, bytecode
and environments are skipped as opaque ``None`` where possible.

Returns a dict {variable name -> numpy array | list | dict | str}.
Matrices keep R's column-major meaning: an R ``T x d`` matrix becomes a numpy
array of shape (T, d).
"""
import gzip
import struct

import numpy as np

NILVALUE_SXP, GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP = 254, 253, 242, 241
REFSXP, NAMESPACESXP, PACKAGESXP, PERSISTSXP = 255, 249, 250, 247
MISSINGARG_SXP, UNBOUNDVALUE_SXP, BASENAMESPACE_SXP = 251, 252, 247
ALTREP_SXP = 238


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def length(self):
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + lo
        return n

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == NILVALUE_SXP:
            return None
        if t in (GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, MISSINGARG_SXP, UNBOUNDVALUE_SXP):
            return None
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            self.i32()
            n = self.i32()
            v = [self.item() for _ in range(n)]
            self.refs.append(None)
            return v
        if t == 4:  # ENVSXP
            self.i32()
            self.refs.append(None)
            for _ in range(4):
                self.item()
            return None
        if t in (2, 6, 5, 17):  # LISTSXP, LANGSXP, PROMSXP, DOTSXP -> python list of (tag, value)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                t2 = flags & 0xFF
                if t2 == NILVALUE_SXP:
                    break
                if t2 not in (2, 6, 5, 17):
                    # cdr is not a pairlist cell; rewind and read as a generic item
                    self.p -= 4
                    out.append((None, self.item()))
                    break
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return out
        if t == 3:  # CLOSXP
            attr = self.item() if has_attr else None
            self.item(); self.item(); self.item()
            return None
        if t == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.p:self.p + n].decode("utf-8", "replace")
            self.p += n
            return s
        if t in (10, 13):  # LGLSXP, INTSXP
            n = self.length()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int32)
            self.p += 4 * n
            res = v
        elif t == 14:  # REALSXP
            n = self.length()
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
            self.p += 8 * n
            res = v
        elif t == 16:  # STRSXP
            n = self.length()
            res = [self.item() for _ in range(n)]
        elif t in (19, 20):  # VECSXP, EXPRSXP
            n = self.length()
            res = [self.item() for _ in range(n)]
        elif t == 24:  # RAWSXP
            n = self.length()
            res = bytes(self.b[self.p:self.p + n])
            self.p += n
        elif t == 21:  # BCODESXP: not needed by any fixture we read
            raise NotImplementedError("bytecode objects are not supported")
        elif t == ALTREP_SXP:
            info = self.item(); state = self.item(); attr = self.item()
            # compact_intseq / compact_realseq: state = (n, start, step)
            cls = info[0][1] if isinstance(info, list) else None
            if cls in ("compact_intseq", "compact_realseq") and state is not None:
                n, start, step = [float(x) for x in np.asarray(state)]
                arr = start + step * np.arange(int(n))
                return arr.astype(np.int32 if cls == "compact_intseq" else np.float64)
            if cls == "wrap_real" or cls == "wrap_integer":
                return state[0] if isinstance(state, list) else state
            raise NotImplementedError(f"ALTREP class {cls}")
        else:
            raise NotImplementedError(f"SEXP type {t} at offset {self.p}")
        if has_attr:
            attrs = dict((k, v) for k, v in self.item())
            if "dim" in attrs and isinstance(res, np.ndarray):
                dim = tuple(int(x) for x in attrs["dim"])
                res = res.reshape(dim, order="F")
            if "names" in attrs and isinstance(res, list):
                res = dict(zip(attrs["names"], res))
            elif "names" in attrs and isinstance(res, np.ndarray) and t == 19:
                res = dict(zip(attrs["names"], res))
        return res


def read_rdata(path):
    """Parse an ``.RData`` file and return {name: value}."""
    with gzip.open(path, "rb") as f:
        buf = f.read()
    if buf[:5] != b"RDX2\n" and buf[:5] != b"RDX3\n":
        raise ValueError("not an RDX2/RDX3 file: %r" % buf[:5])
    version3 = buf[:5] == b"RDX3\n"
    if buf[5:7] != b"X\n":
        raise ValueError("only XDR serialisation is supported")
    r = _Reader(buf)
    r.p = 7
    r.i32(); r.i32(); r.i32()  # format version, R version, min R version
    if version3:
        n = r.i32()
        r.p += n  # native encoding string
    top = r.item()
    return dict((k, v) for k, v in top)
