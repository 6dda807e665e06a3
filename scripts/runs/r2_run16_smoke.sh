# round 2, GPU run 16 (the last seconds of the budget): the final build — smoke(), a slice of the parity tests, one bench line
mkdir -p gpurun_out
timeout 100 python __graft_entry__.py --smoke > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/r2t_smoke.log)"
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or reference_tree or traversal_switches or gather_entry" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2t_pytest.log)"
timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2t_soup.log 2>&1; echo "soup rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2t_soup.log | head -1) $(grep -o '"median": [0-9.]*' gpurun_out/r2t_soup.log | head -1)"
