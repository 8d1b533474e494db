# guarded sweep: parity first, then short benches; every step has a hard timeout
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "k1_ or predict_pipeline or golden_reference" 2>&1 | tail -2 | tee /tmp/sweep_test.log
grep -q "failed\|error" /tmp/sweep_test.log && { echo "PARITY FAILED - not benchmarking"; exit 1; }
RML_FUSED=0 timeout 120 python bench.py --steps 10 --skip-extras 2>/dev/null || echo "bench serial failed/timeout"
for k in ${SWEEP_K2_SMS:-12 20 28}; do RML_K2_SMS=$k timeout 120 python bench.py --steps 10 --skip-extras 2>/dev/null || echo "bench k2_sms=$k failed/timeout"; done
