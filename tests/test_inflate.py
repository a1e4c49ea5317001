"""The host side of the ingest path (ldx_inflate_gz_file): BGZF files -- the format tabix indexes and the reference's
cache holds (prep_intgen_data.py:138) -- are inflated block-parallel, plain and multi-member gzip sequentially; the
text must equal what Python's gzip module produces.  Host code only: runs without a GPU."""
import gzip
import os
import struct
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as g
    g.build()


def bgzf_bytes(data, block=65280, level=6):
    """A BGZF file as htslib writes it: gzip members with a 'BC' extra field holding the member size - 1, then the
    28-byte empty end-of-file block."""
    out = bytearray()
    for a in list(range(0, len(data), block)) + [None]:
        chunk = b"" if a is None else data[a:a + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = c.compress(chunk) + c.flush()
        bsize = 12 + 6 + len(body) + 8
        out += struct.pack("<BBBBIBBH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        out += body + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk))
    return bytes(out)


def vcf_like(n_lines, seed=1):
    rng = np.random.default_rng(seed)
    lines = [b"##fileformat=VCFv4.1", b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + b"\t".join(b"S%d" % i for i in range(40))]
    for k in range(n_lines):
        gt = b"\t".join(b"%d|%d" % (a, b) for a, b in rng.integers(0, 2, size=(40, 2)))
        lines.append(b"22\t%d\trs%d\tA\tG\t100\tPASS\tAC=1;VT=SNP\tGT\t" % (16050000 + 31 * k, 100 + k) + gt)
    return b"\n".join(lines) + b"\n"


@pytest.mark.parametrize("threads", [0, 1, 3])
def test_bgzf_is_inflated_in_parallel_and_equals_gzip(tmp_path, threads):
    from ld_tools_b200 import HostText
    data = vcf_like(6000)                                    # ~1.5 MB of text: a few dozen blocks
    path = tmp_path / "22.vcf.gz"
    path.write_bytes(bgzf_bytes(data))
    assert gzip.decompress(path.read_bytes()) == data        # the writer above makes valid gzip
    t = HostText(str(path), threads=threads)
    assert t.was_bgzf and t.nbytes == len(data) and t.array.tobytes() == data
    t.close()


def test_plain_and_concatenated_gzip(tmp_path):
    from ld_tools_b200 import HostText
    data = vcf_like(3000, seed=2)
    p1 = tmp_path / "plain.gz"
    p1.write_bytes(gzip.compress(data))
    t = HostText(str(p1))
    assert not t.was_bgzf and t.array.tobytes() == data
    p2 = tmp_path / "members.gz"                             # three members without the BC field, one of them empty
    p2.write_bytes(gzip.compress(data[:70000]) + gzip.compress(b"") + gzip.compress(data[70000:]))
    t2 = HostText(str(p2))
    assert not t2.was_bgzf and t2.array.tobytes() == data
    p3 = tmp_path / "tiny.gz"
    p3.write_bytes(gzip.compress(b""))
    assert HostText(str(p3)).nbytes == 0
    big = bytes(np.random.default_rng(0).integers(0, 4, size=40_000_000, dtype=np.uint8))   # expands > 8x: the output buffer grows
    p4 = tmp_path / "big.gz"
    p4.write_bytes(gzip.compress(big, 1))
    assert HostText(str(p4)).array.tobytes() == big


def test_corrupt_files_are_refused(tmp_path):
    from ld_tools_b200 import HostText, LdxError
    data = vcf_like(2000, seed=3)
    good = bytearray(bgzf_bytes(data))
    bad = bytearray(good)
    bad[len(bad) // 2] ^= 0x55                               # a flipped bit inside a block: deflate error or CRC mismatch
    p = tmp_path / "bad.vcf.gz"
    p.write_bytes(bytes(bad))
    with pytest.raises(LdxError):
        HostText(str(p))
    p.write_bytes(gzip.compress(data)[:-9])                  # truncated plain gzip
    with pytest.raises(LdxError):
        HostText(str(p))
    p.write_bytes(b"this is not gzip at all")
    with pytest.raises(LdxError):
        HostText(str(p))
    with pytest.raises(LdxError):
        HostText(str(tmp_path / "missing.gz"))
