# compute-sanitizer memcheck over the kernels changed late in round 2 (K1L forced onto small hostile batches, K2, LZ4 parse/exec, CRC-32 batch)
cd /root/repo
mkdir -p gpurun_out
{
  echo "### SWC_DEFLATE_K1=lut: inflate_lut_kernel + lz_resolve_kernel on ragged / mixed / truncated / malformed / overflowing units"
  SWC_DEFLATE_K1=lut timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_deflate.py -m gpu -x -q -k "ragged or truncation or malformed or overflow or start_bit or inline" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|COMPUTE-SANITIZER" | head -20
  echo "### lz4_parse_kernel / lz4_exec_kernel: ragged + fuzzed blocks, overflow, frames"
  timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_lz4.py -m gpu -x -q -k "ragged or overflow or short or frames" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|COMPUTE-SANITIZER" | head -20
  echo "### crc32_units_kernel (slicing by 4) / xxh32 batches"
  timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_boundary.py -m gpu -x -q -k "checksum" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|COMPUTE-SANITIZER" | head -20
} > gpurun_out/r2_memcheck.txt 2>&1
cat gpurun_out/r2_memcheck.txt
