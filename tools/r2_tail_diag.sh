# 2-GPU box: where the count kernel's tail (grid barrier, peer exchange, ln tables) spends its time; per-rank shards of 25 M
# SNPs (N = 2 of config 5) and of 6.25 M SNPs (the shard size of an 8-GPU run)
mkdir -p gpurun_out
export TDSFS_TAIL_STAMPS=1
for snps in 50000000 12500000; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --snps $snps --steps 50 --warmup 5 --no-cpu --no-e2e --no-extra --verify-windows 0 > gpurun_out/r2t_n2_$snps.json 2> gpurun_out/r2t_n2_$snps.err
  grep "tail stamps" gpurun_out/r2t_n2_$snps.err
  python tools/show_bench.py gpurun_out/r2t_n2_$snps.json
done
