for ctas in 2368 4736 9472; do for ty in 16 32 64; do for nb in 1; do
TAG="ctas=$ctas ty=$ty nb=$nb" AA_TILE_CTAS=$ctas AA_TILE_TY=$ty AA_TILE_NBUF=$nb timeout 60 python scripts/probe_perf.py tile 2>&1 | tail -1
done; done; done
TAG="default" timeout 60 python scripts/probe_perf.py tile 2>&1 | tail -1
