for q in 1024 2048 4096 8192 16384 32768 131072; do echo "Q=$q"; YG_TC_SEG=$q python tools/tc_one_time.py; done
