for f in 0 1; do
  LCF_FLAT=$f python tools/microbench/time_variants.py tools/microbench/variants/timing.so
done
LCF_FLAT=0 python tools/microbench/time_variants.py tools/microbench/variants/head.so tools/microbench/variants/base.so tools/microbench/variants/head.so tools/microbench/variants/base.so
LCF_FLAT=0 python tools/microbench/time_variants.py --walkers 12500 tools/microbench/variants/base.so
LCF_FLAT=1 python tools/microbench/time_variants.py --walkers 12500 tools/microbench/variants/base.so
LCF_FLAT=0 python tools/microbench/time_variants.py --walkers 25000 tools/microbench/variants/base.so
LCF_FLAT=1 python tools/microbench/time_variants.py --walkers 25000 tools/microbench/variants/base.so
