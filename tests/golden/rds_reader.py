"""Minimal reader for R's XDR serialization (``.rds`` / ``.rda``), written for
``make_golden.py`` only.

It exists so that the reference's own golden tables (``R/sysdata.rda`` and
``inst/extdata/*.rds`` under ``/root/reference``) can be turned into small
JSON fixtures in THIS container, where R is not installed.  It is never
imported by the product, the tests or the bench: those read the committed
JSON.

Format notes (R Internals, "Serialization Formats"): every item starts with a
big-endian int32 of flags; the low 8 bits are the SEXPTYPE, 0x100 = is-object,
0x200 = has-attributes, 0x400 = has-tag.  Attributes follow the payload as a
pairlist.  Symbols and environments/external pointers enter a reference
table; type 255 refers back into it.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct
from dataclasses import dataclass, field
from typing import Any

NA_INT = -2147483648


@dataclass
class RObj:
    """Any R value that carries attributes (vectors, lists, S4 instances)."""
    kind: str
    value: Any = None
    attrs: dict = field(default_factory=dict)

    def __getitem__(self, key):
        if isinstance(key, str):
            if self.kind == "list" and "names" in self.attrs:
                names = self.attrs["names"].value
                return self.value[names.index(key)]
            return self.attrs[key]
        return self.value[key]

    def names(self):
        n = self.attrs.get("names")
        return None if n is None else n.value


class _Stop(Exception):
    pass


class Reader:
    def __init__(self, buf: bytes, stop_after_tags: int | None = None):
        self.b = buf
        self.p = 0
        self.refs: list = []
        self.stop_after_tags = stop_after_tags
        self.top_items: list = []

    # -- primitives ---------------------------------------------------------
    def i32(self) -> int:
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def f64s(self, n: int) -> list:
        v = list(struct.unpack_from(">%dd" % n, self.b, self.p))
        self.p += 8 * n
        return v

    def i32s(self, n: int) -> list:
        v = list(struct.unpack_from(">%di" % n, self.b, self.p))
        self.p += 4 * n
        return v

    def length(self) -> int:
        n = self.i32()
        if n == -1:  # long vector
            hi = self.i32()
            lo = self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    # -- items --------------------------------------------------------------
    def item(self, top_level_pairlist: bool = False):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)

        if t == 254:  # NILVALUE_SXP
            return None
        if t == 253:  # GLOBALENV_SXP
            return RObj("env", "global")
        if t == 252:  # UNBOUNDVALUE_SXP
            return RObj("unbound")
        if t == 251:  # MISSINGARG_SXP
            return RObj("missing")
        if t == 250:  # BASENAMESPACE_SXP
            return RObj("env", "basenamespace")
        if t == 242:  # EMPTYENV_SXP
            return RObj("env", "empty")
        if t == 241:  # BASEENV_SXP
            return RObj("env", "base")
        if t == 255:  # reference
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 1:  # symbol
            name = self.item()
            self.refs.append(name)
            return name
        if t in (249, 248, 247):  # NAMESPACESXP / PACKAGESXP / PERSISTSXP: a string vector, entered in the ref table
            self.i32()
            n = self.i32()
            info = [self.item() for _ in range(n)]
            o = RObj("namespace", info)
            self.refs.append(o)
            return o
        if t == 4:  # environment
            o = RObj("env", {})
            self.refs.append(o)
            self.i32()  # locked
            o.value["enclos"] = self.item()
            o.value["frame"] = self.item()
            o.value["hashtab"] = self.item()
            a = self.item()
            if a:
                o.attrs = self._attrs_to_dict(a)
            return o
        if t in (2, 6, 5, 3):  # pairlist, language, promise, closure
            attrs = self.item() if has_attr else None
            out = []
            # iterative CDR walk so long pairlists do not recurse deeply
            while True:
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                if top_level_pairlist:
                    self.top_items.append((tag, car))
                    if self.stop_after_tags and len(self.top_items) >= self.stop_after_tags:
                        raise _Stop()
                # peek at the CDR
                nflags = self.i32()
                nt = nflags & 0xFF
                if nt == 254:
                    break
                if nt not in (2, 6):
                    self.p -= 4
                    out.append(("__cdr__", self.item()))
                    break
                has_tag = bool(nflags & 0x400)
                if nflags & 0x200:
                    self.item()  # attributes of an inner cell: ignored
            o = RObj("pairlist", out)
            if attrs:
                o.attrs = self._attrs_to_dict(attrs)
            return o
        if t == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.p:self.p + n].decode("utf-8", "replace")
            self.p += n
            return s
        if t == 22:  # external pointer
            o = RObj("extptr")
            self.refs.append(o)
            self.item()
            self.item()
            if has_attr:
                self.item()
            return o
        if t == 238:  # ALTREP
            info = self.item()
            state = self.item()
            self.item()  # attributes
            cls = info.value[0][1]
            if cls == "compact_intseq":
                n, start, step = (int(x) for x in state.value)
                return RObj("int", [start + k * step for k in range(n)])
            if cls == "compact_realseq":
                n, start, step = state.value
                return RObj("real", [start + k * step for k in range(int(n))])
            if cls in ("wrap_real", "wrap_integer", "wrap_string", "wrap_logical"):
                return state.value[0][1]
            if cls == "deferred_string":
                inner = state.value[0][1]
                return RObj("str", [_fmt(v) for v in inner.value])
            raise NotImplementedError("ALTREP class %r" % cls)

        # vectors -------------------------------------------------------------
        if t == 10:
            n = self.length()
            o = RObj("lgl", [None if v == NA_INT else bool(v) for v in self.i32s(n)])
        elif t == 13:
            n = self.length()
            o = RObj("int", [None if v == NA_INT else v for v in self.i32s(n)])
        elif t == 14:
            n = self.length()
            o = RObj("real", self.f64s(n))
        elif t == 16:
            n = self.length()
            o = RObj("str", [self.item() for _ in range(n)])
        elif t in (19, 20):
            n = self.length()
            o = RObj("list", [self.item() for _ in range(n)])
        elif t == 24:  # raw
            n = self.length()
            o = RObj("raw", self.b[self.p:self.p + n])
            self.p += n
        elif t == 25:  # S4
            o = RObj("S4")
        elif t == 21:  # byte code (ReadBC in R's serialize.c): parsed only to stay aligned
            nreps = self.i32()
            reps = [None] * nreps
            o = RObj("bytecode", self._bc1(reps))
        elif t in (7, 8):  # special / builtin: a name
            n = self.i32()
            o = RObj("builtin", self.b[self.p:self.p + n].decode())
            self.p += n
        elif t == 15:  # complex
            n = self.length()
            o = RObj("complex", self.f64s(2 * n))
        elif t == 23:  # weak reference
            o = RObj("weakref")
            self.refs.append(o)
        else:
            raise NotImplementedError("SEXPTYPE %d at offset %d" % (t, self.p))
        if has_attr:
            a = self.item()
            if a:
                o.attrs = self._attrs_to_dict(a)
        return o

    # -- byte code ------------------------------------------------------------
    def _bc1(self, reps):
        code = self.item()
        n = self.i32()
        consts = []
        for _ in range(n):
            t = self.i32()
            if t == 21:
                consts.append(self._bc1(reps))
            elif t in (6, 2, 244, 243, 240, 239):
                consts.append(self._bclang(t, reps))
            else:
                consts.append(self.item())
        return (code, consts)

    def _bclang(self, t, reps):
        if t == 243:  # BCREPREF
            return reps[self.i32()]
        if t in (244, 6, 2, 240, 239):
            pos = -1
            if t == 244:  # BCREPDEF
                pos = self.i32()
                t = self.i32()
            hasattr_ = t in (240, 239)
            o = RObj("bclang", None)
            if pos >= 0:
                reps[pos] = o
            attr = self.item() if hasattr_ else None
            tag = self.item()
            car = self._bclang(self.i32(), reps)
            cdr = self._bclang(self.i32(), reps)
            o.value = (attr, tag, car, cdr)
            return o
        return self.item()

    @staticmethod
    def _attrs_to_dict(a: RObj) -> dict:
        return {tag: car for tag, car in a.value if tag is not None}


def _fmt(v):
    if isinstance(v, float) and v == int(v):
        return str(int(v))
    return str(v)


def _decompress(raw: bytes) -> bytes:
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


def _header(r: Reader):
    assert r.b[r.p:r.p + 2] == b"X\n", "only XDR serialization is supported"
    r.p += 2
    version = r.i32()
    r.i32()
    r.i32()
    if version == 3:
        n = r.i32()
        r.p += n


def read_rds(path: str):
    r = Reader(_decompress(open(path, "rb").read()))
    _header(r)
    return r.item()


def read_rda(path: str, n_objects: int):
    """First ``n_objects`` (name, value) pairs of an ``.rda`` file."""
    buf = _decompress(open(path, "rb").read())
    assert buf[:5] == b"RDX2\n" or buf[:5] == b"RDX3\n"
    r = Reader(buf, stop_after_tags=n_objects)
    r.p = 5
    _header(r)
    try:
        r.item(top_level_pairlist=True)
    except _Stop:
        pass
    return r.top_items
