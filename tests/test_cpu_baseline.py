"""The BLAS-backed CPU baseline bench.py times must be the same computation as oracle.c."""
import numpy as np

from oracle import blas_baseline as B
from oracle import oracle as O


def test_blas_baseline_equals_oracle():
    rng = np.random.default_rng(3)
    n, m = 61, 1000
    raw = rng.choice(np.array([0, 1, 2, 3, 6], np.uint8), size=(m, n), p=[.4, .3, .2, .05, .05])
    K, N, done, _ = B.king_blocks(raw, n, m, 300)
    ref = O.snp_king(raw, n, m, block_size_=300)
    assert done == m and np.array_equal(K, ref["K"]) and np.array_equal(N, ref["N_Aa_i"])
    king = B.king_finish(K, N)
    assert np.all((king == ref["king"]) | (np.isnan(king) & np.isnan(ref["king"])))
    K1, _, done1, _ = B.king_blocks(raw, n, m, 300, max_blocks=1)
    assert done1 == int(O.split_len(m, 300)[0, 2])
    assert np.array_equal(K1, O.snp_king(raw, n, m, ind_col=np.arange(1, done1 + 1), block_size_=300)["K"])


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the reference's CPU path restated, on the host cores): one JSON
    line with the keys the driver reads."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pair-SNP/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "50000 x 500000" in line["config"]["workload"] and line["scaling"] == "strong"
