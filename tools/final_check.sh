set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/r02r_bench_default.json 2> gpurun_out/r02r_bench_default.err; tail -c 600 gpurun_out/r02r_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02r_bench_reference.json 2> gpurun_out/r02r_bench_reference.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02r_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_under_ncu.json 2> gpurun_out/r02r_under_ncu.err
ls -la gpurun_out/r02r*
