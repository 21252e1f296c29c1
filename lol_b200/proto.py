"""Wire format of the reference's ring-element messages (lol/Lol.proto:20-49: R, Rq, Kq, RqProduct, KqProduct), so that host
data written by lol / rlwe-challenges (`Protoable (IZipVector m r)`, lol/Crypto/Lol/Types/IZipVector.hs:106-200; `CT` uses the
same instances, CPP.hs:115-121) can be moved to and from the [n][k] device layout without the Haskell side.

proto2 encoding written out by hand (two varint / zigzag helpers; no generated code, no .proto compiler in this image):
    R   { required uint32 m = 1; repeated sint64 xs = 2; }
    Rq  { required uint32 m = 1; required uint64 q = 2; repeated sint64 xs = 3; }
    Kq  { required uint32 m = 1; required uint64 q = 2; repeated double xs = 3; }
    RqProduct { repeated Rq rqlist = 1; }      KqProduct { repeated Kq kqlist = 1; }
Repeated scalars are written unpacked (proto2 default, what hprotoc emits) and read in either form.  The checks on read are
the reference's (`fromProto`: index m, length phi(m), modulus).
"""
from __future__ import annotations

import struct

import numpy as np

from .factored import totient_fact


class ProtoError(ValueError):
    pass


def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf: bytes, pos: int) -> tuple[int, int]:
    v, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise ProtoError("truncated varint")
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v & ((1 << 64) - 1), pos
        shift += 7
        if shift > 63:
            raise ProtoError("varint too long")


def _zigzag(x: int) -> int: return ((x << 1) ^ (x >> 63)) & ((1 << 64) - 1)
def _unzigzag(v: int) -> int: return (v >> 1) ^ -(v & 1)


def _fields(buf: bytes):
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
            if len(v) != ln:
                raise ProtoError("truncated field")
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ProtoError(f"unsupported wire type {wt}")
        yield num, wt, v


def _encode_ring(m: int, q, xs, real: bool) -> bytes:
    out = bytearray(b"\x08" + _varint(int(m)))
    f = 2
    if q is not None:
        out += b"\x10" + _varint(int(q))
        f = 3
    if real:
        key = bytes([(f << 3) | 1])
        for x in np.asarray(xs, dtype=np.float64).ravel():
            out += key + struct.pack("<d", float(x))
    else:
        key = bytes([f << 3])
        for x in np.asarray(xs, dtype=np.int64).ravel():
            out += key + _varint(_zigzag(int(x)))
    return bytes(out)


def _decode_ring(buf: bytes, has_q: bool, real: bool):
    m = q = None
    xs = []
    fx = 3 if has_q else 2
    for num, wt, v in _fields(buf):
        if num == 1 and wt == 0:
            m = v
        elif has_q and num == 2 and wt == 0:
            q = v
        elif num == fx:
            if real:
                if wt == 1:
                    xs.append(struct.unpack("<d", v)[0])
                elif wt == 2:      # packed
                    xs.extend(struct.unpack(f"<{len(v) // 8}d", v))
                else:
                    raise ProtoError("bad wire type for double")
            else:
                if wt == 0:
                    xs.append(_unzigzag(v))
                elif wt == 2:      # packed
                    p = 0
                    while p < len(v):
                        u, p = _read_varint(v, p)
                        xs.append(_unzigzag(u))
                else:
                    raise ProtoError("bad wire type for sint64")
    if m is None or (has_q and q is None):
        raise ProtoError("missing required field")
    arr = np.array(xs, dtype=np.float64 if real else np.int64)
    if arr.size != totient_fact(int(m)):
        raise ProtoError(f"Expected n={totient_fact(int(m))}, got {arr.size}")      # IZipVector.hs:121-123
    return int(m), (int(q) if has_q else None), arr


def encode_R(m: int, xs) -> bytes: return _encode_ring(m, None, xs, False)
def decode_R(buf: bytes): m, _, xs = _decode_ring(buf, False, False); return m, xs
def encode_Rq(m: int, q: int, xs) -> bytes: return _encode_ring(m, q, xs, False)
def decode_Rq(buf: bytes): return _decode_ring(buf, True, False)
def encode_Kq(m: int, q: int, xs) -> bytes: return _encode_ring(m, q, xs, True)
def decode_Kq(buf: bytes): return _decode_ring(buf, True, True)


def encode_RqProduct(m: int, qs, y) -> bytes:
    """y: [n, k] residues (the device / ABI layout of one ring element) -> RqProduct with one Rq per limb, first limb first
    (the tuple instances of IZipVector.hs concatenate the lists in component order)."""
    y = np.asarray(y, dtype=np.int64).reshape(-1, len(qs))
    out = bytearray()
    for t, q in enumerate(qs):
        body = encode_Rq(m, q, y[:, t])
        out += b"\x0a" + _varint(len(body)) + body
    return bytes(out)


def decode_RqProduct(buf: bytes, m: int | None = None, qs=None) -> tuple[int, list[int], np.ndarray]:
    """-> (m, qs, y [n, k]).  With m / qs given, checks them like `fromProto` does (IZipVector.hs:141-162)."""
    limbs = [decode_Rq(v) for num, wt, v in _fields(buf) if num == 1 and wt == 2]
    if not limbs:
        raise ProtoError("empty RqProduct")
    ms = {l[0] for l in limbs}
    if len(ms) != 1 or (m is not None and ms != {int(m)}):
        raise ProtoError(f"Expected m={m}, got {sorted(ms)}")
    got = [l[1] for l in limbs]
    if qs is not None and [int(q) for q in qs] != got:
        raise ProtoError(f"Expected q={list(qs)}, got {got}")
    y = np.stack([np.mod(l[2], l[1]) for l in limbs], axis=1)      # `reduce` on read (IZipVector.hs:162)
    return limbs[0][0], got, np.ascontiguousarray(y)


def encode_KqProduct(m: int, qs, y) -> bytes:
    y = np.asarray(y, dtype=np.float64).reshape(-1, len(qs))
    out = bytearray()
    for t, q in enumerate(qs):
        body = encode_Kq(m, q, y[:, t])
        out += b"\x0a" + _varint(len(body)) + body
    return bytes(out)


def decode_KqProduct(buf: bytes) -> tuple[int, list[int], np.ndarray]:
    limbs = [decode_Kq(v) for num, wt, v in _fields(buf) if num == 1 and wt == 2]
    if not limbs or len({l[0] for l in limbs}) != 1:
        raise ProtoError("bad KqProduct")
    return limbs[0][0], [l[1] for l in limbs], np.ascontiguousarray(np.stack([l[2] for l in limbs], axis=1))
