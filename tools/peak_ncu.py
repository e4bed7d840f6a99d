import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pspace_b200.device import fp64_peak
print("dmma", fp64_peak(1, 50.0), "dfma", fp64_peak(0, 50.0))
