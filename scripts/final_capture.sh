set -x
D=gpurun_out/${1:-r2f}; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; tail -2 $D/pytest.log
python bench.py > $D/bench_n1.json 2> $D/bench_n1.err
python bench.py --impl reference > $D/bench_ref.json 2> $D/bench_ref.err
for w in readme pdbbind misato misato_production stress; do python bench.py --workload $w > $D/bench_$w.json 2> $D/bench_$w.err; done
python scripts/profile_step.py 2 > $D/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/launches.csv python bench.py --steps 2 --warmup 1 > $D/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -c 6 -o $D/prof_full -f python scripts/profile_step.py 1 > $D/ncu2.log 2>&1
ncu -i $D/prof_full.ncu-rep --page raw --csv > $D/prof_full_raw.csv 2>/dev/null
python scripts/profile_k3.py > $D/k3_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:aggregate -c 5 -o $D/prof_k3 -f python scripts/profile_k3.py > $D/ncu3.log 2>&1
ncu -i $D/prof_k3.ncu-rep --page raw --csv > $D/prof_k3_raw.csv 2>/dev/null
python scripts/bench_kernels.py > $D/kernel_roofline.json 2> $D/bk.err
python scripts/bench_calls.py > $D/calls.json 2> $D/calls.err
for w in "dynamic 4" "readme 4" "pdbbind 64" "misato 4" "stress 1" "residue_single 1"; do python scripts/time_workload.py $w; done > $D/time_workloads.jsonl 2> $D/tw.err
python scripts/k3_pipe_probe.py > $D/k3_pipe.txt 2>&1
rm -f $D/prof_full.ncu-rep.tmp; ls -la $D | head -40
