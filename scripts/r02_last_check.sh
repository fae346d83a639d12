timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or mosaic" 2>&1 | tail -2; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
