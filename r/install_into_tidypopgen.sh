#!/bin/bash
# Copies the B200 backend into a tidypopgen source tree so that `R CMD INSTALL` builds it through
# src/Makevars (nvcc rules, sm_100a), as BASELINE.json's north_star asks:
#     r/install_into_tidypopgen.sh /path/to/tidypopgen && R CMD INSTALL /path/to/tidypopgen
# What changes in the package: src/ gains the backend sources + gpu_glue.cpp and its Makevars is
# replaced; R/snp_king.R, R/snp_ibs.R and R/snp_allele_sharing.R get the one-call bodies of
# r/R/snp_king_gpu.R; Rcpp::compileAttributes() regenerates the .Call registration.
set -euo pipefail
PKG="${1:?usage: install_into_tidypopgen.sh <tidypopgen source dir>}"
HERE="$(cd "$(dirname "$0")/.." && pwd)"
test -f "$PKG/DESCRIPTION" && grep -q '^Package: tidypopgen' "$PKG/DESCRIPTION" || { echo "$PKG is not a tidypopgen source tree" >&2; exit 1; }
CS="$HERE/tidypopgen_b200/csrc"
# sources: the includes are rewritten from the repository layout ("../../include/x.h") to a flat src/
for f in api.cu decode.cu gram.cu finish.cu relate.cu wg17.cu peak.cu pack.cpp config.cpp internal.h config.h hostpool.h ptx.cuh swar.cuh ddsum.cuh; do
  sed 's#"\.\./\.\./include/#"#' "$CS/$f" > "$PKG/src/$f"
done
cp "$HERE/include/tpg_b200.h" "$HERE/include/tpg_synth.h" "$PKG/src/"
cp "$HERE/r/src/gpu_glue.cpp" "$PKG/src/gpu_glue.cpp"
cp "$HERE/r/src/Makevars" "$PKG/src/Makevars"
cp "$HERE/r/R/snp_king_gpu.R" "$PKG/R/snp_gpu_drivers.R"
# the three reference drivers are superseded by R/snp_gpu_drivers.R (same names, same signatures)
for f in snp_king.R snp_ibs.R snp_allele_sharing.R; do
  [ -f "$PKG/R/$f" ] && mv "$PKG/R/$f" "$PKG/R/$f.cpu_reference.txt"
done
if command -v Rscript >/dev/null 2>&1; then
  Rscript -e "Rcpp::compileAttributes('$PKG')"
else
  echo "Rscript not found: run Rcpp::compileAttributes('$PKG') before R CMD INSTALL" >&2
fi
echo "backend installed into $PKG (nvcc: ${CUDA_HOME:-/usr/local/cuda}/bin/nvcc)"
