for st in 4 3 2; do echo "stages=$st"; SPN_DEBUG_STAGES=$st EXP_VARIANTS=tp,hg6 timeout 300 python tools/exp_fused.py C3 | grep case; done
