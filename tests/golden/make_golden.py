"""Builds tests/golden/* from the reference's own fixtures (run once in the dev container,
where /root/reference exists; the GPU box only sees the committed outputs).

    python tests/golden/make_golden.py

Inputs (read-only, reference data files, not source):
  inst/extdata/related/families.{bed,fam,bim}   12 x 961 PLINK bed with 446 NAs
      (built by data-raw/create_ibs_dataset.R:48-79, set.seed(123))
  inst/extdata/related/test_king.kin0           output of `king -b families_k.bed --kinship`
      (used by tests/testthat/test_snp_king.R:84-131: cols 2,3 = pair, col 8 = kinship)
  inst/extdata/related/test_plinkIBS.mibs       output of `plink --distance square flat-missing ibs`
      (used by tests/testthat/test_snp_ibs.R:69-104)
  inst/extdata/bigsnp_obj_gt/old_gen_tibble.bk  the 7 x 6 example FBM (load_example_gt())
Outputs:
  families.bk           FBM.code256 bytes, column-major 12 x 961, NA = 3, as
                        bigsnpr::snp_readBed / gen_tibble(valid_alleles = c("1","2")) write them
  families_kin0.json    [[id1, id2, kinship], ...]  (66 pairs)
  families_mibs.json    12 x 12 PLINK IBS proportions (6 decimals)
  example_gt.bk         42 bytes, column-major 7 x 6
  golden_meta.json      shapes + sha256 of every file
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/inst/extdata")
OUT = Path(__file__).resolve().parent


def read_bed(path: Path, n: int, m: int) -> np.ndarray:
    """PLINK .bed (SNP-major) -> bigSNP bytes [m, n].

    2-bit codes, lowest bits first: 00 -> 2 copies of A1, 01 -> missing, 10 -> het,
    11 -> 0 copies; bigsnpr's getCode() maps them to 2, NA(3), 1, 0.  All four statistics are
    invariant to swapping 0 <-> 2, so the allele orientation cannot affect parity."""
    raw = np.frombuffer(path.read_bytes(), dtype=np.uint8)
    assert raw[0] == 0x6C and raw[1] == 0x1B and raw[2] == 0x01, "not a SNP-major .bed"
    bpl = (n + 3) // 4
    body = raw[3:].reshape(m, bpl)
    code = np.array([2, 3, 1, 0], dtype=np.uint8)
    out = np.zeros((m, n), np.uint8)
    for i in range(n):
        out[:, i] = code[(body[:, i // 4] >> (2 * (i % 4))) & 3]
    return out


def main():
    if not REF.exists():
        sys.exit("reference not mounted; the committed fixtures are the product of this script")
    n = len((REF / "related/families.fam").read_text().strip().splitlines())
    m = len((REF / "related/families.bim").read_text().strip().splitlines())
    assert (n, m) == (12, 961)
    g = read_bed(REF / "related/families.bed", n, m)  # [m, n] == column-major n x m
    assert int((g == 3).sum()) == 446, "the reference fixture has 446 missing genotypes"
    (OUT / "families.bk").write_bytes(g.tobytes())

    kin = []
    for ln in (REF / "related/test_king.kin0").read_text().strip().splitlines():
        f = ln.split()
        kin.append([int(f[1]), int(f[2]), float(f[7])])
    assert len(kin) == 66
    (OUT / "families_kin0.json").write_text(json.dumps(kin))

    mibs = [[float(x) for x in ln.split()] for ln in
            (REF / "related/test_plinkIBS.mibs").read_text().strip().splitlines()]
    assert len(mibs) == 12 and all(len(r) == 12 for r in mibs)
    (OUT / "families_mibs.json").write_text(json.dumps(mibs))

    ex = (REF / "bigsnp_obj_gt/old_gen_tibble.bk").read_bytes()
    assert len(ex) == 42
    (OUT / "example_gt.bk").write_bytes(ex)

    meta = {"families": {"n": n, "m": m, "n_missing": 446}, "example_gt": {"n": 7, "m": 6}, "sha256": {}}
    for f in ["families.bk", "families_kin0.json", "families_mibs.json", "example_gt.bk"]:
        meta["sha256"][f] = hashlib.sha256((OUT / f).read_bytes()).hexdigest()
    (OUT / "golden_meta.json").write_text(json.dumps(meta, indent=1))
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
