"""lol_b200/proto.py against the protobuf runtime itself: the messages of lol/Lol.proto:20-49 are declared to the official
Python runtime through a FileDescriptorProto built here (no protoc in the image), then bytes are exchanged both ways."""
import numpy as np
import pytest

from lol_b200 import proto

pb = pytest.importorskip("google.protobuf")
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory      # noqa: E402

F = descriptor_pb2.FieldDescriptorProto


def _messages():
    fd = descriptor_pb2.FileDescriptorProto(name="lol_wire_test.proto", package="crypto.proto.lol.test", syntax="proto2")

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = tname
    msg("R", [("m", 1, F.TYPE_UINT32, F.LABEL_REQUIRED, None), ("xs", 2, F.TYPE_SINT64, F.LABEL_REPEATED, None)])
    msg("Rq", [("m", 1, F.TYPE_UINT32, F.LABEL_REQUIRED, None), ("q", 2, F.TYPE_UINT64, F.LABEL_REQUIRED, None),
               ("xs", 3, F.TYPE_SINT64, F.LABEL_REPEATED, None)])
    msg("Kq", [("m", 1, F.TYPE_UINT32, F.LABEL_REQUIRED, None), ("q", 2, F.TYPE_UINT64, F.LABEL_REQUIRED, None),
               ("xs", 3, F.TYPE_DOUBLE, F.LABEL_REPEATED, None)])
    msg("RqProduct", [("rqlist", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".crypto.proto.lol.test.Rq")])
    msg("KqProduct", [("kqlist", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".crypto.proto.lol.test.Kq")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    return {n: (get(pool.FindMessageTypeByName("crypto.proto.lol.test." + n)) if get else
                message_factory.MessageFactory(pool).GetPrototype(pool.FindMessageTypeByName("crypto.proto.lol.test." + n)))
            for n in ("R", "Rq", "Kq", "RqProduct", "KqProduct")}


def test_wire_format_matches_the_protobuf_runtime():
    M = _messages()
    rng = np.random.default_rng(3)
    m, n = 21, 12
    xs = rng.integers(-2 ** 62, 2 ** 62, size=n)
    xs[:3] = [0, -1, 2 ** 63 - 1]
    ref = M["R"](m=m, xs=[int(v) for v in xs])
    assert proto.encode_R(m, xs) == ref.SerializeToString()
    mm, back = proto.decode_R(ref.SerializeToString())
    assert mm == m and np.array_equal(back, xs)

    qs = [19393921, 18869761]
    y = np.stack([rng.integers(0, q, size=n) for q in qs], axis=1)
    refp = M["RqProduct"](rqlist=[M["Rq"](m=m, q=q, xs=[int(v) for v in y[:, t]]) for t, q in enumerate(qs)])
    assert proto.encode_RqProduct(m, qs, y) == refp.SerializeToString()
    m2, q2, y2 = proto.decode_RqProduct(refp.SerializeToString(), m, qs)
    assert (m2, q2) == (m, qs) and np.array_equal(y2, y)
    parsed = M["RqProduct"].FromString(proto.encode_RqProduct(m, qs, y))
    assert [list(r.xs) for r in parsed.rqlist] == [[int(v) for v in y[:, t]] for t in range(2)]

    k = rng.normal(size=(n, 1))
    refk = M["KqProduct"](kqlist=[M["Kq"](m=m, q=qs[0], xs=[float(v) for v in k[:, 0]])])
    assert proto.encode_KqProduct(m, qs[:1], k) == refk.SerializeToString()
    assert np.array_equal(proto.decode_KqProduct(refk.SerializeToString())[2], k)


def test_reads_packed_fields_and_applies_the_reference_checks():
    m, n = 7, 6
    xs = np.arange(-3, 3)
    body = b"".join(proto._varint(proto._zigzag(int(v))) for v in xs)
    packed = b"\x08" + proto._varint(m) + b"\x12" + proto._varint(len(body)) + body
    assert np.array_equal(proto.decode_R(packed)[1], xs)
    with pytest.raises(proto.ProtoError):      # wrong length (IZipVector.hs:121-123)
        proto.decode_R(proto.encode_R(m, xs[:5]))
    with pytest.raises(proto.ProtoError):      # wrong modulus (IZipVector.hs:156-158)
        proto.decode_RqProduct(proto.encode_RqProduct(m, [29], xs % 29), m, [43])
    # negative representatives are reduced on read, like `reduce` in fromProto
    assert np.array_equal(proto.decode_RqProduct(proto.encode_RqProduct(m, [29], xs))[2][:, 0], xs % 29)
