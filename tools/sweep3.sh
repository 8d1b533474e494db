for sp in 1 2 4; do RML_K1_SPLIT=$sp timeout 120 python bench.py --steps 10 --skip-extras 2>/dev/null | cut -c60-260; done
