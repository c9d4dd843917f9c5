#!/bin/bash
T=${1:-r02t}
mkdir -p gpurun_out
cp lighter_b200/liblighter_b200.so /tmp/default.so
python tools/march_variant_diff.py dump config4 /tmp/fv_hint.npz
cp lighter_b200/variants/lib_mh0.so lighter_b200/liblighter_b200.so
python tools/march_variant_diff.py dump config4 /tmp/fv_plain.npz
cp /tmp/default.so lighter_b200/liblighter_b200.so
python tools/march_variant_diff.py check config4 /tmp/fv_plain.npz /tmp/fv_hint.npz > gpurun_out/${T}_march_diff_config4.log 2>&1
tail -50 gpurun_out/${T}_march_diff_config4.log
