for k in 24 32 40 48 64; do RML_K2_SMS=$k timeout 120 python bench.py --steps 10 --skip-extras 2>/dev/null | cut -c60-300 || echo "bench k2_sms=$k failed/timeout"; done
