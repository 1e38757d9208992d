#!/bin/bash
# Other BASELINE configs for the record: config 1 (B=1 audio only), config 3 (A+V+F, bf16 / bf16x3)
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/cfg_$tag.json 2> gpurun_out/cfg_$tag.err; python -c "
import json; d=json.load(open('gpurun_out/cfg_$tag.json')); print('$tag', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['roofline']['frac'],4), d['roofline']['breakdown_ms_per_step'])"; }
run c1_audio_b1 --batch 1 --encoders audio
run c3_avf_bf16 --encoders audio,video,flow --precision bf16
run c3_avf_bf16x3 --encoders audio,video,flow --precision bf16x3
run c2_av_bf16 --encoders audio,video --precision bf16
