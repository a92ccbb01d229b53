import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import umma_micro as m
c, s = int(sys.argv[1]), int(sys.argv[2])
m.bench(c, s, iters=3)
