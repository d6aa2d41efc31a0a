import cProfile, pstats, sys, warnings
sys.path.insert(0, '/root/repo')
import bench
from lightcurve_fitting_b200 import synthetic
from lightcurve_fitting_b200.bolometric import calculate_bolometric
lc = synthetic.sed_table(bench.device_truth, 500, seed=2)
warnings.simplefilter('ignore')
calculate_bolometric(lc.copy(), res=1., seed=2)
pr = cProfile.Profile()
pr.enable()
t, tm = calculate_bolometric(lc.copy(), res=1., seed=3, return_timing=True)
pr.disable()
print(tm)
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
