timeout 900 python bench.py --mode train --steps 10 --warmup 3 2>&1 | tail -3 | cut -c1-3000
