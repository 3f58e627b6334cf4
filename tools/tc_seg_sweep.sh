for v in 0 64 128 256; do echo "L2PROMO=$v"; YG_TC_L2PROMO=$v python tools/tc_one_time.py; done
