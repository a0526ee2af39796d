#!/bin/bash
# Does the UNMODIFIED reference fastgs CUDA build run at the BASELINE sizes on this box?  One thing varied at a time
# (VERDICT r1 item 1).  Writes everything under gpurun_out/fastgs_exp/.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/fastgs_exp
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > $O/box.txt 2>&1
python tools/dump_scene.py C2 /tmp/c2.bin
python tools/dump_scene.py C3 /tmp/c3.bin
for var in static shared nofast sm100; do
  for cfg in c2 c3; do
    timeout 300 oracle/_ref/fastgs_standalone_$var /tmp/$cfg.bin --check --views 2 > $O/standalone_${var}_${cfg}.jsonl 2> $O/standalone_${var}_${cfg}.err
    echo "exit $?" >> $O/standalone_${var}_${cfg}.jsonl
  done
done
# through the reference's own forward_wrapper inside a torch process
for cfg in C2 C3; do
  timeout 600 python tools/ref_train.py --module ref --config $cfg --check > $O/torch_ref_${cfg}_check.json 2> $O/torch_ref_${cfg}_check.err
  echo "exit $?" >> $O/torch_ref_${cfg}_check.json
done
# sanitizer passes on the smallest failing size (C2), standalone static build
for tool in memcheck initcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 40 oracle/_ref/fastgs_standalone_static /tmp/c2.bin --check --views 1 \
     > $O/sanitizer_${tool}_c2.txt 2>&1
  echo "exit $?" >> $O/sanitizer_${tool}_c2.txt
done
# timings where it runs
for var in static; do
  timeout 600 oracle/_ref/fastgs_standalone_$var /tmp/c3.bin --check --train --views 8 --steps 3 --warmup 1 > $O/standalone_${var}_c3_train.jsonl 2> $O/standalone_${var}_c3_train.err
  echo "exit $?" >> $O/standalone_${var}_c3_train.jsonl
done
timeout 900 python tools/ref_train.py --module ref --config C3 --views 8 --steps 3 --warmup 1 > $O/torch_ref_C3_train.json 2> $O/torch_ref_C3_train.err
echo "exit $?" >> $O/torch_ref_C3_train.json
timeout 900 python tools/ref_train.py --module b200 --config C3 --views 8 --steps 3 --warmup 1 > $O/torch_b200_C3_train.json 2> $O/torch_b200_C3_train.err
echo "exit $?" >> $O/torch_b200_C3_train.json
tail -n 3 $O/*.jsonl $O/*.json | cut -c1-600
