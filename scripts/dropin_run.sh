#!/bin/bash
# SURVEY 8(f)-1 acceptance run at C1's size: the reference's sample driver (src/samples/test_spmv.c), UNMODIFIED,
# linked against the reference library (CPU, all host threads) and against libspmv_b200.so (GPU), on the same
# Matrix-Market file.  Both print the reference's CSV schema; host x / y in both cases.
set -e
mkdir -p gpurun_out /tmp/dropin/mtx_cache
python - <<'PY'
import sys
sys.path.insert(0, ".")
from spmv_b200 import matrices as M, mtx
mtx.write_mtx("/tmp/dropin/lap1024.mtx", M.laplacian2d(1024), symmetric=True)
PY
T=$(nproc)
cd /tmp/dropin
echo "# nproc=$T; columns: matrix,method,vectorized,threads,nnz,err,pre_ms,avg_ms,GFLOPS_avg,GFLOPS_best" > $OLDPWD/gpurun_out/dropin_c1.csv
echo "# --- reference library (CPU) ---" >> $OLDPWD/gpurun_out/dropin_c1.csv
$OLDPWD/oracle/_ref/test_spmv_ref lap1024.mtx $T $T >> $OLDPWD/gpurun_out/dropin_c1.csv
echo "# --- libspmv_b200.so (B200), default settings (the driver's recurring X / Y are page-locked in place) ---" >> $OLDPWD/gpurun_out/dropin_c1.csv
$OLDPWD/oracle/_ref/test_spmv_b200 lap1024.mtx $T $T >> $OLDPWD/gpurun_out/dropin_c1.csv
echo "# --- libspmv_b200.so (B200), SPMV_B200_PIN_HOST=0: pageable X / Y staged by the CUDA driver ---" >> $OLDPWD/gpurun_out/dropin_c1.csv
SPMV_B200_PIN_HOST=0 $OLDPWD/oracle/_ref/test_spmv_b200 lap1024.mtx $T $T >> $OLDPWD/gpurun_out/dropin_c1.csv
cat $OLDPWD/gpurun_out/dropin_c1.csv
