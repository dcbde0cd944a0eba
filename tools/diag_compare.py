import sys
import numpy as np
base = np.load(sys.argv[1])
for p in sys.argv[2:]:
    o = np.load(p)
    print(p, {k: int((base[k].view(np.uint32) != o[k].view(np.uint32)).sum()) for k in base.files})
