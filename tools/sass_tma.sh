#!/bin/bash
# usage: tools/sass_tma.sh > profiles/sass_rNN_tma.txt -- the TMA / mbarrier instructions of the two kernels that stage tiles by TMA and the static
# mnemonic census of the FAST kernel, from cuobjdump -sass of the built library (no GPU needed)
cd "$(dirname "$0")/.."
SO=gtsam_vslam_b200/libgvs.so
body() { cuobjdump -sass $SO 2>/dev/null | awk -v k="$1" '/Function :/{f=index($0,k)>0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/\/\*[0-9a-f]\{16\}\*\///; s/ *$//'; }
echo "# SASS evidence for the TMA tile staging: cuobjdump -sass $SO, every UTMALDG / mbarrier (SYNCS) / ELECT / async fence of the two"
echo "# kernels that stage by TMA (line number inside the kernel : address : instruction), plus the static mnemonic census of"
echo "# fast_cells_kernel<true>.  Regenerate: tools/sass_tma.sh"
for k in fast_cells_kernelILb1 pyr_fused_kernel; do
    echo; echo "## $k"
    body $k | grep -n -E "UTMALDG|SYNCS|ELECT|FENCE.VIEW" | head -40
done
echo; echo "## mnemonic census of fast_cells_kernel<true> (static instruction count)"
body fast_cells_kernelILb1 | awk '{m=$2; if (m ~ /^@/) m=$3; sub(/\..*/,"",m); c[m]++} END{for (k in c) print c[k], k}' | sort -rn | head -30
