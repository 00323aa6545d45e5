"""(f3) .cmt wire format: the library's bincode writer / reader against a HAND-DERIVED bincode 1.3 encoding of
`ReefCommitment` (commitment.rs:44-52, merkle_tree.rs:10-15, main.rs:37-51).  Host-only code: runs in the CPU tier."""
import ctypes as C
import random
import struct

import numpy as np
import pytest

import reef_b200
from oracle.fields import FQ
from oracle.merkle import MerkleCommitment as OracleMerkle
from reef_b200._lib import ReefError, check, lib


def fe(x):
    return int(x).to_bytes(32, "little")


def hand_bincode_merkle(commitment, tree, doc, orig_len):
    """bincode 1.3 default options: u64 little-endian lengths, u8 Option tags, struct fields in declaration order,
    field elements as bare 32-byte arrays"""
    out = b"\x00"                                         # nldoc: None
    out += b"\x01"                                        # merkle: Some(MerkleCommitment {
    out += fe(commitment)                                 #   commitment: F,
    out += struct.pack("<Q", len(tree))                   #   tree: Vec<Vec<F>>,
    for level in tree:
        out += struct.pack("<Q", len(level)) + b"".join(fe(x) for x in level)
    out += struct.pack("<Q", len(doc)) + b"".join(fe(x) for x in doc)    # doc: Vec<F> })
    out += struct.pack("<Q", orig_len)                    # orig_doc_len: usize
    out += struct.pack("<Q", len(doc))                    # udoc_len: usize
    return out


@pytest.mark.parametrize("n", [1, 2, 5, 8, 33])
def test_merkle_cmt_bytes_and_round_trip(n):
    rnd = random.Random(n)
    doc = [rnd.randrange(130) for _ in range(n)]
    mc = OracleMerkle(doc)
    exp = hand_bincode_merkle(mc.commitment, mc.tree, doc, n - 1)
    sizes = np.asarray([len(l) for l in mc.tree], dtype=np.uint64)
    levels = b"".join(fe(x) for l in mc.tree for x in l)
    d = np.asarray(doc, dtype=np.uint64)
    size = lib.reef_cmt_merkle_size(sizes.ctypes.data, len(sizes), n)
    assert size == len(exp)
    out, out_len = C.create_string_buffer(size), C.c_uint64()
    check(lib.reef_cmt_merkle_write(fe(mc.commitment), levels, sizes.ctypes.data, len(sizes), d.ctypes.data, n, n - 1, out, size, C.byref(out_len)))
    assert out.raw[:out_len.value] == exp
    with pytest.raises(ReefError):
        lib_check_small = lib.reef_cmt_merkle_write(fe(mc.commitment), levels, sizes.ctypes.data, len(sizes), d.ctypes.data, n, n - 1, out, size - 1,
                                                    C.byref(out_len))
        check(lib_check_small)
    # read back: sizing call, then the data
    kind = C.c_int(-1)
    check(lib.reef_cmt_probe(exp, len(exp), C.byref(kind)))
    assert kind.value == 1
    nl, nn, dl, ol, ul = C.c_uint32(), C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    check(lib.reef_cmt_merkle_read(exp, len(exp), None, None, 0, None, 0, C.byref(nl), C.byref(nn), None, 0, C.byref(dl), C.byref(ol), C.byref(ul)))
    assert (nl.value, nn.value, dl.value, ol.value, ul.value) == (len(mc.tree), sum(len(l) for l in mc.tree), n, n - 1, n)
    com, lv = C.create_string_buffer(32), C.create_string_buffer(32 * nn.value)
    ls, dd = np.zeros(nl.value, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
    check(lib.reef_cmt_merkle_read(exp, len(exp), com, lv, nn.value, ls.ctypes.data, nl.value, C.byref(nl), C.byref(nn), dd.ctypes.data, n,
                                   C.byref(dl), C.byref(ol), C.byref(ul)))
    assert com.raw == fe(mc.commitment) and lv.raw == levels and list(ls) == list(sizes) and list(dd) == doc
    # the reference `expect("Could not deserialize")`s on a truncated / over-long file
    for bad in (exp[:-1], exp + b"\x00", b"\x02" + exp[1:], exp[:40]):
        with pytest.raises(ReefError) as e:
            check(lib.reef_cmt_merkle_read(bad, len(bad), None, None, 0, None, 0, C.byref(nl), C.byref(nn), None, 0, C.byref(dl), C.byref(ol), C.byref(ul)))
        assert e.value.code == 3


def test_point_compression_is_x_with_the_parity_of_y():
    import workloads as WL
    pts = WL.generators("pallas", 8)
    for k in range(8):
        x, y = int.from_bytes(pts[64 * k:64 * k + 32], "little"), int.from_bytes(pts[64 * k + 32:64 * k + 64], "little")
        out = C.create_string_buffer(32)
        check(lib.reef_point_compress(pts[64 * k:64 * k + 64], out))
        assert int.from_bytes(out.raw, "little") == x | ((y & 1) << 255)
    out = C.create_string_buffer(32)
    check(lib.reef_point_compress(bytes(64), out))
    assert out.raw == bytes(32)


def test_nldoc_cmt_bytes():
    """NLDocCommitment (commitment.rs:54-68): Reef-computed members encoded here, nova-snark-owned members opaque"""
    from reef_b200._lib import CmtNldoc
    import workloads as WL
    rnd = random.Random(3)
    num_vars, rows = 4, 4
    codes = np.asarray([rnd.randrange(130) for _ in range(11)], dtype=np.uint32)
    pts = WL.generators("pallas", rows)
    pts = pts[:64] + bytes(64) + pts[128:]                       # one identity commitment
    blinds = b"".join(fe(rnd.randrange(FQ)) for _ in range(rows))
    h, salt = fe(rnd.randrange(FQ)), fe(rnd.randrange(FQ))
    sg, hg, pk, vk = b"SINGLE-GENS", b"HYRAX-GEN-BYTES", b"PK", b"VERIFIER-KEY"
    f = CmtNldoc()
    keep = [C.create_string_buffer(x, len(x)) for x in (sg, hg, pts, blinds, h, salt, pk, vk)]
    f.single_gens, f.single_gens_len = C.addressof(keep[0]), len(sg)
    f.hyrax_gen, f.hyrax_gen_len = C.addressof(keep[1]), len(hg)
    f.num_vars, f.doc_codes, f.doc_len = num_vars, codes.ctypes.data, len(codes)
    f.row_commitments, f.blinds, f.rows = C.addressof(keep[2]), C.addressof(keep[3]), rows
    f.doc_commit_hash, f.hash_salt = C.addressof(keep[4]), C.addressof(keep[5])
    f.cap_pk, f.cap_pk_len, f.cap_vk, f.cap_vk_len = C.addressof(keep[6]), len(pk), C.addressof(keep[7]), len(vk)
    f.q_len, f.orig_doc_len, f.udoc_len = num_vars, 9, 16
    exp = b"\x01" + sg + hg
    exp += struct.pack("<Q", num_vars) + struct.pack("<Q", 16) + b"".join(fe(int(c)) for c in codes) + bytes(32 * 5)
    comp = b""
    for k in range(rows):
        x, y = int.from_bytes(pts[64 * k:64 * k + 32], "little"), int.from_bytes(pts[64 * k + 32:64 * k + 64], "little")
        comp += (x | ((y & 1) << 255)).to_bytes(32, "little")
    exp += struct.pack("<Q", rows) + comp + struct.pack("<Q", rows) + blinds + h + salt + pk + vk + struct.pack("<Q", num_vars)
    exp += b"\x00" + struct.pack("<QQ", 9, 16)
    size = lib.reef_cmt_nldoc_size(C.byref(f))
    assert size == len(exp)
    out, out_len = C.create_string_buffer(size), C.c_uint64()
    check(lib.reef_cmt_nldoc_write(C.byref(f), out, size, C.byref(out_len)))
    assert out.raw == exp
    kind = C.c_int(-1)
    check(lib.reef_cmt_probe(out.raw, size, C.byref(kind)))
    assert kind.value == 0
