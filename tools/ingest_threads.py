#!/usr/bin/env python
"""Reader-only thread sweep of the C++ block VCF reader (open / read_block / parse seconds, best of 3) on one file.

    python tools/ingest_threads.py FILE [KEY ...]        # default keys: DP DFLANKINDEL Q
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trtools_b200.vcf_ingest import NativeVCF      # noqa: E402


def main():
    path = sys.argv[1]
    keys = tuple(sys.argv[2:]) or ("DP", "DFLANKINDEL", "Q")
    size = os.path.getsize(path)
    print("# %s (%.1f MB on disk), keys GT+%s, host threads %d" % (os.path.basename(path), size / 1e6, "+".join(keys),
                                                                  os.cpu_count()))
    print("threads  open_s  read_s  parse_s  total_s  loci  MB_text  MB/s_text")
    t = 1
    while t <= (os.cpu_count() or 1):
        best = None
        for _ in range(3):
            t0 = time.time()
            v = NativeVCF(path, threads=t)
            v._readahead = False
            v._native_block_loci = 1 << 30
            v._native_block_bytes = 0
            t1 = time.time()
            v._next_block()
            t2 = time.time()
            v._blk.parse(keys)
            t3 = time.time()
            r = (t1 - t0, t2 - t1, t3 - t2, v._blk.n, int(v._blk.line_off[-1]))
            if best is None or sum(r[:3]) < sum(best[:3]):
                best = r
            v.close()
        tot = sum(best[:3])
        print("%7d  %.3f   %.3f   %.3f    %.3f    %d  %.1f   %.0f" % (t, best[0], best[1], best[2], tot, best[3],
                                                                     best[4] / 1e6, best[4] / 1e6 / tot))
        t *= 2


if __name__ == "__main__":
    main()
