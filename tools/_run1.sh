mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_env.txt 2>&1; nproc >> gpurun_out/r2a_env.txt
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2a_pytest.log
(timeout 1200 bash tools/variants_ntt.sh "-DSTARK_FIELD_CARRY=0" "" "-DSTARK_NTT_REGCAP=48" "-DSTARK_CORR_FMA=0" "-DSTARK_CORR_FMA=0 -DSTARK_NTT_REGCAP=48" "-DSTARK_CORR_FMA=1" "-DSTARK_CORR_FMA=6" "-DSTARK_CORR_FMA=3 -DSTARK_NTT_REGCAP=48" 2>&1) > gpurun_out/r2a_variants.txt
tail -5 gpurun_out/r2a_pytest.log
