for v in 0 1 3; do K4LZ4_COPY_VARIANT=$v python tools/dbench.py --blocks 65536 2>&1 | tail -1; done
