cd /root/repo
bash scripts/ab_lib.sh glob glob2 glob glob2 glob glob2
PB200_LIB_PATH=$PWD/build/variants/glob2.so timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
