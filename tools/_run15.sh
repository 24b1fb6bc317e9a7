set -x
mkdir -p gpurun_out
N=$1
if [ "$N" = "1" ]; then
  timeout 900 python tools/sweep_cfg5.py --frames 100000 > gpurun_out/r2_cfg5_n1.json 2> gpurun_out/r2_cfg5_n1.err || tail -5 gpurun_out/r2_cfg5_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N tools/sweep_cfg5.py --frames 100000 > gpurun_out/r2_cfg5_n$N.json 2> gpurun_out/r2_cfg5_n$N.err || tail -5 gpurun_out/r2_cfg5_n$N.err
fi
cat gpurun_out/r2_cfg5_n$N.json
