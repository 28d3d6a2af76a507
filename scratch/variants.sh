timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for v in 0 2; do K4LZ4_COPY_VARIANT=$v python tools/dbench.py --blocks 65536 2>&1 | tail -1; done
