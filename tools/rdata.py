"""Minimal reader for R's XDR serialisation (RDX2 / RDX3 workspaces), enough for GRanges fixtures.

Used ONLY to generate committed fixtures (tools/make_golden.py) from the reference's shipped
data/ExomeCount.RData; nothing at test/run time reads /root/reference.  Written from the format
description in R-ints ("Serialization Formats"); no R code involved.
"""
import gzip
import struct


class RObj:
    __slots__ = ("kind", "value", "attr", "tag")

    def __init__(self, kind, value=None, attr=None, tag=None):
        self.kind, self.value, self.attr, self.tag = kind, value, attr or {}, tag

    def __repr__(self):
        v = self.value
        if isinstance(v, (list, tuple)) and len(v) > 6:
            v = f"<{len(v)} items>"
        return f"RObj({self.kind}, {v}, attr={list(self.attr)})"


class Reader:
    def __init__(self, buf):
        self.b, self.p, self.refs = buf, 0, []

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

    def raw(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def attrs(self):
        out = {}
        o = self.item()
        while o is not None and o.kind == "pairlist":
            for tag, val in o.value:
                out[tag] = val
            break
        return out

    def pairlist(self, has_attr, has_tag, first_flags=None):
        items = []
        attr = {}
        flags_has_attr, flags_has_tag = has_attr, has_tag
        while True:
            if flags_has_attr:
                attr = self.attrs()
            tag = None
            if flags_has_tag:
                t = self.item()
                tag = t.value if t is not None else None
            car = self.item()
            items.append((tag, car))
            # cdr
            flags = self.i32()
            ty = flags & 0xFF
            if ty == 254:  # NILVALUE
                break
            if ty != 2:
                # improper list tail; parse as item with these flags
                self.p -= 4
                items.append((None, self.item()))
                break
            flags_has_attr = bool(flags & (1 << 9))
            flags_has_tag = bool(flags & (1 << 10))
        return RObj("pairlist", items, attr)

    def item(self):
        flags = self.i32()
        ty = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if ty == 254:
            return None
        if ty in (253, 252, 251, 242, 241):  # global/empty/base env, missing arg, unbound
            return RObj("special", ty)
        if ty == 255:
            return self.refs[(flags >> 8) - 1] if (flags >> 8) else self.refs[self.i32() - 1]
        if ty == 1:  # SYMSXP
            name = self.item()
            o = RObj("symbol", name.value)
            self.refs.append(o)
            return o
        if ty in (249, 250, 247):  # namespace / package / persist
            self.i32()
            n = self.i32()
            vals = [self.item() for _ in range(n)]
            o = RObj("namespace", vals)
            self.refs.append(o)
            return o
        if ty == 4:  # ENVSXP
            o = RObj("env", {})
            self.refs.append(o)
            self.i32()  # locked
            enclos, frame, hashtab, attr = self.item(), self.item(), self.item(), self.item()
            o.value = {"frame": frame, "hashtab": hashtab}
            return o
        if ty in (2, 6, 5, 17, 239, 240):  # LISTSXP, LANGSXP, PROMSXP, DOTSXP, ATTRLIST/ATTRLANG
            if ty in (2, 239):
                return self.pairlist(has_attr, has_tag)
            attr = self.attrs() if has_attr else {}
            tag = self.item() if has_tag else None
            car, cdr = self.item(), self.item()
            return RObj("lang", (car, cdr), attr)
        if ty == 3:  # CLOSXP
            attr = self.attrs() if has_attr else {}
            env, formals, body = self.item(), self.item(), self.item()
            return RObj("closure", None, attr)
        if ty == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return RObj("char", None)
            return RObj("char", self.raw(n).decode("utf-8", "replace"))
        if ty == 10 or ty == 13:
            n = self.length()
            v = list(struct.unpack_from(f">{n}i", self.b, self.p))
            self.p += 4 * n
            o = RObj("lgl" if ty == 10 else "int", v)
        elif ty == 14:
            n = self.length()
            v = list(struct.unpack_from(f">{n}d", self.b, self.p))
            self.p += 8 * n
            o = RObj("real", v)
        elif ty == 16:
            n = self.length()
            o = RObj("str", [self.item().value for _ in range(n)])
        elif ty in (19, 20):
            n = self.length()
            o = RObj("list", [self.item() for _ in range(n)])
        elif ty == 24:
            n = self.length()
            o = RObj("raw", self.raw(n))
        elif ty == 25:
            o = RObj("S4")
        elif ty == 238:  # ALTREP
            info, state, attr = self.item(), self.item(), self.item()
            return RObj("altrep", (info, state))
        else:
            raise ValueError(f"unsupported SEXP type {ty} at offset {self.p}")
        if has_attr:
            o.attr = self.attrs()
        return o


def load(path):
    """Return {name: RObj} for an .RData workspace."""
    with gzip.open(path, "rb") as fh:
        buf = fh.read()
    assert buf[:5] in (b"RDX2\n", b"RDX3\n"), buf[:5]
    r = Reader(buf)
    r.p = 5
    assert r.raw(2) == b"X\n"
    version = r.i32()
    r.i32(); r.i32()
    if version == 3:
        n = r.i32()
        r.raw(n)
    top = r.item()
    return {tag: val for tag, val in top.value}
