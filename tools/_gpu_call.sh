set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k fastgs 2>&1 | tail -30
timeout 300 python tests/golden/make_fastgs_golden.py gpurun_out/fastgs_ref_golden.npz 2>&1 | tail -3
timeout 300 python - <<'PY' 2>&1 | tail -30
import sys, json
sys.path.insert(0, 'tests')
import gpu_diag as D
for kw in (dict(), dict(n=1500, w=120, h=100, deg=1, seed=5, sigma_px=7.0)):
    try:
        r = D.diag_fastgs(**kw)
        print(json.dumps({k: (round(v, 8) if isinstance(v, float) else v) for k, v in r.items()}))
    except Exception as e:
        import traceback; traceback.print_exc()
PY
