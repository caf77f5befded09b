for f in _ab/*.so; do B200MC_SO=$PWD/$f timeout 120 python tools/ab_models.py 2>&1 | tail -1; done
