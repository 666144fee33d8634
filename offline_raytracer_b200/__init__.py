"""offline_raytracer_b200 -- B200 (sm_100a) hot path of gyuhyun-lee/offline_raytracer.

The product is libort_b200.so (C ABI in include/ort_b200.h, sources in csrc/);
this package is the thin host-side binding used by the tests and the benchmark.
"""
from .api import (  # noqa: F401
    ORT_KERNEL_DEFAULT, ORT_KERNEL_MEGAKERNEL, ORT_KERNEL_WAVEFRONT, MISS_RANK,
    OrtError, Camera, RenderParams, RenderStats, Scene, HostScene,
    default_params, device_count, lib, load_mesh, parse_numeric, write_hdr, v3_to_rgbe,
    measure_fp32_peak, measure_l2_bandwidth, selftest_div3, selftest_intersect, selftest_bsdf, selftest_rng, selftest_light_pick, world_light_is_sphere,
    generate_camera_rays_device, generate_random_rays_device, Multi, Progress, bake_mesh, LIB_PATH, ORT_BUILD_ON_DEVICE, ORT_HOST_NO_OCTREE, ShapeLists,
)
