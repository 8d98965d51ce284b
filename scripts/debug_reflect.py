import sys, numpy as np
sys.path.insert(0, '.')
from tests import common
from oracle import bindings
from raytracercpp_b200 import api
lib = api.load_library()
z = np.load('tests/golden/robot_scene.npz'); rows = z['materials']
mats = [dict(ambient_coeff=tuple(r[0:3]), diffuse=tuple(r[3:6]), specular=tuple(r[6:9]), emission=tuple(r[9:12]), reflection=float(r[12]), roughness=float(r[13]), ns=float(r[14]), specular_threshold=float(r[15])) for r in rows]
robot = dict(xyz9=z['xyz9'], uv6=z['uv6'], mat=z['mat'])
orc = bindings.CpuTracer('oracle')
kw0, m, tex = common.config_table(mats)['cfg3_mirror5']
for depth in (0,1,2,3,5):
  for shadows in (0,1):
    kw = dict(kw0, max_recursion_depth=depth, compute_shadows=shadows)
    img, st = common.product_image(lib, robot, kw, m, tex)
    r = common.oracle_renderer(orc, robot, kw, m, tex); want = r.render()[0]; cnt = r.count_rows()
    print(depth, shadows, 'gpu', st.reflection_rays, st.reflection_shadow_rays, 'oracle', cnt['reflection_rays'], cnt['reflection_shadow_rays'], 'img', common.image_error(img, want), (img!=want).sum())
