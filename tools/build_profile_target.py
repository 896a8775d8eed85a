"""ncu target: one production build of a scene (default: the 10,000,200-triangle instanced soup)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanort_b200 import api, scenes as S
v, f = S.make_scene(sys.argv[1] if len(sys.argv) > 1 else "instanced")
acc = api.BVHAccel()
acc.Build(len(f), v, f)
st = acc.GetStatistics()
print(f"built {len(f)} tris: device {st['build_secs']*1e3:.2f} ms")
