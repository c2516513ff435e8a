#!/bin/bash
# 2-GPU session: thread transport across two devices, NCCL transport under torchrun, bench --gpus 2
set -u
OUT=gpurun_out/s5
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt; nvidia-smi topo -m >> $OUT/gpus.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > $OUT/pytest_sharded.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_sharded.log
tail -30 $OUT/pytest_sharded.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741"
timeout 600 $TR tools/sharded_check.py > $OUT/sharded_check.log 2>&1; echo "check rc=$?" >> $OUT/sharded_check.log
grep -E "OK|MISMATCH|SHARDED_CHECK|rc=|Error|error" $OUT/sharded_check.log | head -30
timeout 600 $TR bench.py --gpus 2 --workload genome100m --steps 3 --warmup 3 > $OUT/bench2_g100m.json 2> $OUT/bench2_g100m.err; echo "rc=$?"
cat $OUT/bench2_g100m.json; tail -5 $OUT/bench2_g100m.err
timeout 1200 $TR bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench2_genome3g.json 2> $OUT/bench2_genome3g.err; echo "rc=$?"
cat $OUT/bench2_genome3g.json; tail -5 $OUT/bench2_genome3g.err
timeout 600 $TR bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > $OUT/bench2_reference.json 2>&1
cat $OUT/bench2_reference.json
