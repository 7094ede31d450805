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


# ---------------------------------------------------------------------------------------------- writer
class _Writer:
    """Serialises python values in R's XDR format (version 2): the inverse of _Reader for the subset the reference's scripts
    exchange -- numeric / integer / logical / character vectors and matrices, named lists, NULL, Inf and NA."""

    def __init__(self):
        self.out = bytearray()
        self.syms = {}

    def i32(self, v):
        self.out += struct.pack(">i", int(v))

    def flags(self, t, has_attr=False, has_tag=False, is_obj=False):
        self.i32(t | (int(is_obj) << 8) | (int(has_attr) << 9) | (int(has_tag) << 10))

    def charsxp(self, s):
        if s is None:
            self.i32(9)                      # NA_STRING
            self.i32(-1)
            return
        b = s.encode("utf-8")
        self.i32(9 | (1 << 18 if b.isascii() else 1 << 15))   # levels << 12: ASCII_MASK (1 << 6) or UTF8_MASK (1 << 3)
        self.i32(len(b))
        self.out += b

    def symbol(self, name):
        if name in self.syms:
            self.i32((self.syms[name] << 8) | REFSXP)
            return
        self.i32(1)
        self.charsxp(name)
        self.syms[name] = len(self.syms) + 1

    def attributes(self, attrs):
        for k, v in attrs:
            self.flags(2, has_tag=True)
            self.symbol(k)
            self.item(v)
        self.i32(NILVALUE_SXP)

    def item(self, v):
        if v is None:
            self.i32(NILVALUE_SXP)
            return
        if isinstance(v, dict):
            self.flags(19, has_attr=True)
            self.i32(len(v))
            for x in v.values():
                self.item(x)
            self.attributes([("names", [str(k) for k in v.keys()])])
            return
        if isinstance(v, (list, tuple)) and v and all(isinstance(x, (str, type(None))) for x in v):
            self.flags(16)
            self.i32(len(v))
            for x in v:
                self.charsxp(x)
            return
        if isinstance(v, (list, tuple)):
            self.flags(19)
            self.i32(len(v))
            for x in v:
                self.item(x)
            return
        if isinstance(v, str):
            self.item([v])
            return
        a = np.asarray(v)
        attrs = []
        if a.ndim >= 2:
            attrs.append(("dim", np.asarray(a.shape, dtype=np.int32)))
        flat = np.ravel(a, order="F")
        if a.dtype == np.bool_:
            self.flags(10, has_attr=bool(attrs))
            self.i32(flat.size)
            self.out += flat.astype(">i4").tobytes()
        elif np.issubdtype(a.dtype, np.integer):
            self.flags(13, has_attr=bool(attrs))
            self.i32(flat.size)
            self.out += flat.astype(">i4").tobytes()
        else:
            self.flags(14, has_attr=bool(attrs))
            self.i32(flat.size)
            self.out += flat.astype(">f8").tobytes()
        if attrs:
            self.attributes(attrs)


def write_rdata(path, variables):
    """Write {name: value} as a gzip'd RDX2 file that R's load() reads (and read_rdata() reads back): numpy arrays become
    numeric / integer / logical vectors or matrices (column-major, dim attribute), dicts named lists, lists unnamed lists
    or character vectors, None NULL.  This is how c_chains lists (samples1, samples2, log_pdf1, log_pdf2, meetingtime,
    iteration, finished -- R/debiasing_algorithm.R:254-263) and data sets travel between this package and the R scripts."""
    w = _Writer()
    w.out += b"RDX2\nX\n"
    w.i32(2)                # serialisation format version
    w.i32(0x00030603)       # written by "R 3.6.3"
    w.i32(0x00020300)       # readable from R 2.3.0
    for name, val in variables.items():
        w.flags(2, has_tag=True)
        w.symbol(str(name))
        w.item(val)
    w.i32(NILVALUE_SXP)
    with gzip.open(path, "wb") as f:
        f.write(bytes(w.out))
