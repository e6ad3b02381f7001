#!/bin/bash
# Round-2 evidence: GPU tests, smoke, bench (both arms), cfg timings, ncu launch list, traffic, ncu --set full of the top kernels
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/gpu_tests.txt 2>&1; echo "tests rc=$?"; tail -3 $O/gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.txt
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 600 python profiles/cfg_timings.py > $O/cfg_timings.json 2> $O/cfg_timings.err; echo "cfg rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 > $O/ncu_launches.log 2>&1; echo "launches rc=$?"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:duo_kernel -s 3 -c 1 --csv --log-file $O/traffic_full_size.csv python bench.py --steps 2 --warmup 3 --skip-configs > $O/ncu_traffic.log 2>&1; echo "traffic rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:relay -s 4 -c 1 -o $O/relay_cfg3 -f python profiles/tools/r02_one.py cfg3 16384 > $O/ncu_relay.log 2>&1; echo "relay ncu rc=$?"
timeout 300 ncu --metrics sm__inst_executed_pipe_fp32.sum,smsp__issue_active.sum,sm__cycles_elapsed.sum,smsp__inst_executed.sum --clock-control none -k regex:relay -s 4 -c 1 --csv --log-file $O/relay_cfg3_fp32.csv python profiles/tools/r02_one.py cfg3 16384 > $O/ncu_relay_fp32.log 2>&1; echo "relay fp32 rc=$?"
ls -la $O | tail -20
