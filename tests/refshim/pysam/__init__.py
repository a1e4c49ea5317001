"""TEST INFRASTRUCTURE ONLY: a pure-Python stand-in for the slice of pysam (htslib) that ld-tools'
drivers use, so that the UNMODIFIED reference drivers (ld_area.py, ld_triangle.py, ld_lite.py) can
run in a container without pysam and serve as end-to-end oracles (tests/golden/make_driver_golden.py).

Surface: VariantFile(path) as a context manager; .fetch() and
.fetch(chrom, start, end) with htslib's 0-based half-open OVERLAP semantics (a record is returned
iff POS-1 < end and record_end > start; record_end = POS-1 + len(REF), or the value of INFO/END when the record
has one greater than POS-1); record attributes id, chrom, pos (1-based),
ref, alts (tuple), info (mapping; flags are keys; VT -> tuple of str), samples[name]['GT'] -> tuple of
ints / None; tabix_index() is a no-op.  Nothing in ld_tools_b200/ imports this package.

PROVENANCE.  Real pysam / htslib / tabix are not in this image and cannot be installed (no network), so these
semantics are restated from htslib's published sources, not observed: the interval of a VCF line in a tabix index
and in the iterator's per-line re-check is tbx.c:tbx_parse1, VCF preset -- begin = POS-1, end = begin + len(REF)
(column 4), replaced by INFO's END=<n> when column 8 starts with "END=" or contains ";END=" and n > begin (an END
at or before begin is warned about and ignored); the iterator returns a line iff end > region_start and begin <
region_end.  GT decoding follows pysam's VariantRecordSample: alleles as ints, '.' -> None, one entry per allele of
the call (so a haploid call is a 1-tuple), phasing separator '|' or '/'.  INFO flags are keys of rec.info.  Golden
cases that pin the edge semantics the engine hard-codes: tests/golden/drivers/area_edge_* (an indel straddling the
window's left edge, one ending exactly at it, records at high-1 / high, END= first and in the middle of INFO, END
equal to the window start, CIEND= without END).  A maintainer with pysam installed can regenerate the same cases
with tests/golden/make_driver_golden.py after removing tests/refshim from PYTHONPATH.
"""
import gzip


class _Sample(dict):
    pass


class _Samples:
    def __init__(self, names_to_col, fields):
        self._cols, self._fields = names_to_col, fields

    def __getitem__(self, name):
        gt = self._fields[self._cols[name]]           # KeyError for an unknown sample, as in pysam
        sep = "|" if "|" in gt else "/"
        return _Sample(GT=tuple(None if a == "." else int(a) for a in gt.split(":")[0].split(sep)))


class _Record:
    def __init__(self, line, names_to_col):
        f = line.rstrip("\n").split("\t")
        self.chrom, self.pos, self.id, self.ref = f[0], int(f[1]), f[2], f[3]
        self.alts = tuple(f[4].split(","))
        self.info = {}
        for item in f[7].split(";"):
            if "=" in item:
                k, v = item.split("=", 1)
                self.info[k] = tuple(v.split(",")) if k == "VT" else v
            elif item:
                self.info[item] = True
        self.samples = _Samples(names_to_col, f[9:])
        self.start, self.stop = self.pos - 1, self.pos - 1 + len(self.ref)
        # tbx.c:tbx_parse1 (VCF preset): "END=" at the start of INFO, else ";END="
        s = f[7]
        k = 4 if s.startswith("END=") else (s.find(";END=") + 5 if ";END=" in s else -1)
        if k >= 0:
            digits = ""
            while k < len(s) and s[k].isdigit():
                digits += s[k]
                k += 1
            if digits and int(digits) > self.start:
                self.stop = int(digits)


class VariantFile:
    def __init__(self, path, mode="r"):
        self._recs, names = [], []
        with gzip.open(path, "rt") as fh:
            for line in fh:
                if line.startswith("##"):
                    continue
                if line.startswith("#"):
                    names = line.rstrip("\n").split("\t")[9:]
                    cols = {n: i for i, n in enumerate(names)}
                    continue
                self._recs.append(_Record(line, cols))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    def fetch(self, contig=None, start=None, stop=None):
        for r in self._recs:
            if contig is not None and r.chrom != contig:
                continue
            if start is not None and not (r.start < stop and r.stop > start):
                continue
            yield r


def tabix_index(*args, **kwargs):
    return None
