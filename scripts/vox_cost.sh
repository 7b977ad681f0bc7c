for p in 4 2 1; do for s in 5 3 2; do echo "place_per_sm=$p scatter_per_sm=$s"; PP_PLACE_PER_SM=$p PP_SCATTER_PER_SM=$s python scripts/inflight.py 12 2>&1 | grep -E "stages=v |stages=ven"; done; done
PP_PLACE_PER_SM=2 PP_SCATTER_PER_SM=3 python scripts/inflight.py 1 2>&1 | grep -E "stages=v |stages=ven"
