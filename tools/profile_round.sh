#!/bin/sh
# Round profile on a GPU box: (1) plain bench, (2) ncu launch list of the same command, (3) ncu --set full of the dominant
# kernel in both launch forms (one launch each).  Everything lands in gpurun_out/; tools/ncu_summary.py turns the reports
# into profiles/*.md + *_traffic.json here.
R=${1:-r02}
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_plain.json 2> gpurun_out/${R}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_list.log 2>&1
# serial instantiation (what roofline.frac times): the pipeline-depth-1 pass of perf_k6a.py
ncu --set full --import-source on --clock-control none -k regex:k_fri_merkle -s 0 -c 1 -f -o gpurun_out/${R}_k_fri_merkle_serial \
    python tools/perf_k6a.py 16384 > gpurun_out/${R}_ncu_full_serial.log 2>&1
# pipelined instantiation (what the timed region launches): skip the two serial launches; two chunks of 8192 proofs, so that the
# call takes the multi-lane path (one chunk would run the serial instantiation again)
P2V_PERF_CHUNK=8192 ncu --set full --import-source on --clock-control none -k regex:k_fri_merkle -s 2 -c 1 -f -o gpurun_out/${R}_k_fri_merkle_pipe \
    python tools/perf_k6a.py 16384 > gpurun_out/${R}_ncu_full_pipe.log 2>&1
tail -2 gpurun_out/${R}_ncu_full_serial.log gpurun_out/${R}_ncu_full_pipe.log
