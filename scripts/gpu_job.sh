mkdir -p gpurun_out
echo "=== memcheck: tensor-core GEMMs (tf32 + bf16 flavours) and one training step on each chain"
timeout -s KILL 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "test_tf32_tensor_core_gemm_vs_torch or test_bf16_storage_tensor_core_gemm_vs_torch or (test_training_step_tensor_core_vs_fp32_path and kw2)" 2>&1 | tail -6
