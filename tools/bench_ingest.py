"""Host-ingest throughput: libvf_ingest.so vs the pure-Python restatement (oracle/ingest_py.py) on seeded synthetic
files (FASTA in BGZF form, multi-sample VCF in BGZF form).  CPU only; writes profiles/r01_ingest.json."""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import ingest_py  # noqa: E402
from tests.test_ingest_native import _bgzf  # noqa: E402
from variantformer_b200 import ingest  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seq = rng.choice(np.frombuffer(b"ACGTacgtN", np.uint8), mb << 20).tobytes()
    fasta = b">chr1\n" + b"\n".join(seq[i:i + 60] for i in range(0, len(seq), 60)) + b"\n"
    n_rec = 1_000_000
    pos = np.sort(rng.integers(1, mb << 20, n_rec))
    b = np.array([b"A", b"C", b"G", b"T"])
    ref, alt = b[rng.integers(0, 4, n_rec)], b[rng.integers(0, 4, n_rec)]
    gts = np.array([b"0/0", b"0/1", b"1/1", b"1|0"])
    g = [gts[rng.integers(0, 4, n_rec)] for _ in range(8)]
    lines = [b"##fileformat=VCFv4.2", b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + b"\t".join(b"S%d" % i for i in range(8))]
    lines += [b"chr1\t%d\t.\t%s\t%s\t.\tPASS\t.\tGT\t%s" % (p, r, a, b"\t".join(x[i] for x in g))
              for i, (p, r, a) in enumerate(zip(pos.tolist(), ref.tolist(), alt.tolist()))]
    vcf = b"\n".join(lines) + b"\n"
    res = {"cores": os.cpu_count(), "fasta_bytes": len(fasta), "vcf_bytes": len(vcf), "vcf_records": n_rec}
    with tempfile.TemporaryDirectory() as d:
        fa, vc = os.path.join(d, "g.fa.gz"), os.path.join(d, "s.vcf.gz")
        open(fa, "wb").write(_bgzf(fasta)); open(vc, "wb").write(_bgzf(vcf))
        for name, fn_n, fn_p, nbytes in (("fasta", lambda: ingest.load_fasta(fa), lambda: ingest_py.load_fasta(fa), len(fasta)),
                                         ("vcf", lambda: ingest.load_vcf_sample(vc, sample="S3"),
                                          lambda: ingest_py.load_vcf_sample(vc, sample="S3"), len(vcf))):
            t0 = time.perf_counter(); a = fn_n(); t1 = time.perf_counter(); bpy = fn_p(); t2 = time.perf_counter()
            res[name] = {"native_s": t1 - t0, "python_s": t2 - t1, "native_MBps": nbytes / (t1 - t0) / 1e6,
                         "python_MBps": nbytes / (t2 - t1) / 1e6, "speedup": (t2 - t1) / (t1 - t0)}
            if name == "fasta":
                assert np.array_equal(a["chr1"], bpy["chr1"])
            else:
                assert np.array_equal(a["chr1"]["pos"], bpy["chr1"]["pos"]) and a["chr1"]["alt"] == bpy["chr1"]["alt"]
    print(json.dumps(res))
    json.dump(res, open("profiles/r01_ingest.json", "w"), indent=1)


if __name__ == "__main__":
    main()
