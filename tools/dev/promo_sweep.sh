for v in 128 64 0 256; do
  echo "== SSIM_CUDA_L2_PROMOTION=$v"
  SSIM_CUDA_L2_PROMOTION=$v timeout 300 python tools/dev/regimes.py quick 2>&1 | grep -E "4K pair|strip|64 x 4K"
  SSIM_CUDA_L2_PROMOTION=$v timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ssim_fused -s 2 -c 1 python tools/dev/prof_batch.py 64 3 2>&1 | grep -E "dram__bytes|gpu__time"
done
