cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; tail -c 600 gpurun_out/bench_r01_final.json; tail -3 gpurun_out/bench_r01_final.err
python bench.py --impl reference > gpurun_out/bench_r01_final_ref.json 2>> gpurun_out/bench_r01_final.err; cat gpurun_out/bench_r01_final_ref.json | cut -c1-400
for E in 1024 4096 131072; do
  ncu --set full --clock-control none --import-source on -k regex:d2d_step_warp -s 2 -c 1 -f -o gpurun_out/prof_final_E$E python profiles/prof_step.py $E 4 > gpurun_out/prof_final_E$E.log 2>&1
  python profiles/ncu_summary.py gpurun_out/prof_final_E$E.ncu-rep $E > gpurun_out/ncu_final_E$E.txt 2>&1
  [ $E = 4096 ] || rm -f gpurun_out/prof_final_E$E.ncu-rep
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:d2d_ --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 64 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/launches_r01c.csv
export D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/timeline.so
for E in 1024 4096; do timeout 100 python profiles/timeline.py $E > gpurun_out/timeline_final_E$E.txt 2>&1; rm -f gpurun_out/timeline_E$E.npy; done
tail -12 gpurun_out/timeline_final_E1024.txt
