mkdir -p gpurun_out
echo "=== all gpu tests"; timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error|assert" | head -20
echo "=== sanitizer memcheck on the fp32 + streaming kernels (tensor-core kernel excluded: tcgen05 under memcheck is very slow)"
timeout -s KILL 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "raw2outputs_shapes or sample_pdf_bit_exact or merge_sorted or zvals or train_mode_and_gradients[render_train_small] or fused_kde or fused_adam or rays_from_pose" 2>&1 | tail -6
