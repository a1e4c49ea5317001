"""TEST INFRASTRUCTURE ONLY: a pure-Python stand-in for the slice of pysam (htslib) that ld-tools'
drivers use, so that the UNMODIFIED reference drivers (ld_area.py, ld_triangle.py, ld_lite.py) can
run in a container without pysam and serve as end-to-end oracles (tests/golden/make_driver_golden.py).

Surface: VariantFile(path) as a context manager; .fetch() and
.fetch(chrom, start, end) with htslib's 0-based half-open OVERLAP semantics (a record is returned
iff POS-1 < end and POS-1+rlen > start, rlen = len(REF)); record attributes id, chrom, pos (1-based),
ref, alts (tuple), info (mapping; flags are keys; VT -> tuple of str), samples[name]['GT'] -> tuple of
ints / None; tabix_index() is a no-op.  Nothing in ld_tools_b200/ imports this package.
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
