rm -f gpurun_out/baseline_parity.jsonl
timeout 1000 python -m pytest tests/test_baseline_configs_gpu.py tests/test_effdet_gpu.py::test_d7_full_size_matches_oracle tests/test_multi_camera_p2p_gpu.py -q --timeout=600 2>&1 | tail -60
echo "=== margins"; cat gpurun_out/baseline_parity.jsonl
bash tools/round2_experiments.sh
