"""raytracercpp_b200 -- B200-native (sm_100a) implementation of the ray-tracing hot path of TomClabault/RayTracerCPP.

The product is `librtb200.so` (CUDA kernels + host octree builder behind the C ABI of include/rtb200.h); this
package is its ctypes binding plus a Python mirror of the reference's `Renderer` interface.  No CPU fallback.
"""
from .api import (Context, RtError, RtMaterial, RtRenderStats, RtSettings, default_settings, load_library,  # noqa: F401
                  RT_TEX_AO, RT_TEX_DIFFUSE, RT_TEX_NORMAL, RT_TEX_ROUGHNESS, RT_TEX_SKYSPHERE)
from .renderer import Renderer, precompute_materials, render  # noqa: F401
