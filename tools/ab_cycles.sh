#!/bin/bash
# Same-box A/B of library variants: tools/ab_cycles.sh tmp_ab/libA.so tmp_ab/libB.so ...
# For each variant: elapsed SM cycles + duration of the LSTM kernels (ncu, one pass, no clock control) at the bench shape.
# Cycle counts drift a few % between boxes (SM / HBM clock ratio), so only numbers from ONE call are comparable.
for rep in 1 2; do
for lib in "$@"; do
    cp "$lib" generative-audio_b200/libnppc_b200.so
    out=gpurun_out/ab_$(basename "$lib" .so)_$rep.csv
    timeout 150 ncu --metrics sm__cycles_elapsed.max,gpu__time_duration.sum --clock-control none -k regex:"lstm_rec|gemm_zx" -c 6 \
        --csv --log-file "$out" python tools/lstm_bench.py 64 ${AB_O:-10} > /dev/null 2>&1
    echo "== $lib rep $rep"
    grep -v "^==" "$out" | tail -6 | awk -F'","' '{n=$5; sub(/\(.*/,"",n); printf "%s %s %s\n", substr(n,1,40), $(NF-2), $NF}' | tr -d '"'
done
done
