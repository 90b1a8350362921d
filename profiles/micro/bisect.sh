for i in 1 2 3 4 5 6 7 8; do
  r=$(timeout 150 python -m pytest tests/test_gpu_graph.py tests/test_gpu_parity.py -q -x -k "(adam_device or embed_gather) or gemm_variants" 2>&1 | grep -E "passed|failed" | tail -1)
  echo "run $i: $r"
done
for i in 1 2 3 4; do timeout 150 python -m pytest tests -m gpu -q -x 2>&1 | tail -1; done
