"""The bounds-checked build of the kernels (csrc/rt_types.h: RT_CHECKED; include/rt_b200.h: rt_checked_build, rt_violations) run over
the parity tests: the GPU counterpart of the reference's sanitizer options (framework/cmake/Sanitizers.cmake:7-37).  The pool these
kernels are developed on has no compute-sanitizer, so the library carries its own: every index formed from data (BVH nodes, triangles,
stack slots, 8-wide nodes and group stacks, accumulator / framebuffer pixels, texels, material / light / sphere tables) is tested
against its bound and violations are counted per site.  The same tests must pass (the checks change no result) and count nothing."""
import json
import os
import subprocess
import sys

import pytest

from util import ROOT

CHECKED_LIB = os.path.join(ROOT, "raytracer-group27_b200", "librtb200_checked.so")
P, Q = "tests/test_gpu_parity.py", "tests/test_gpu_post.py"
# every kernel family at fixture size: three builders, per-level / whole-path / eight-lanes-per-ray traversal, all light kinds,
# transparent shadows, spheres, textures with the mip filters, multi-sample frames, overflow retry, gather frames, post-processing
SUBSET = [
    f"{P}::test_golden[ploc-cornell_c1_256]", f"{P}::test_golden[lbvh-cornell_c1_256]", f"{P}::test_golden[sah-cornell_ms16_70x45]",
    f"{P}::test_golden[ploc-cornell_c4_96]", f"{P}::test_golden[ploc-cornell_planelight_160]", f"{P}::test_golden[ploc-cube_preset_spot_128]",
    f"{P}::test_golden[ploc-tex_trilinear_repeat_clamp_96x80]", f"{P}::test_golden[ploc-tex_mipnearest_floor64_96x80]",
    f"{P}::test_golden[ploc-tex_bilinear_clamp_repeat_96x80]", f"{P}::test_golden[ploc-texdebug_bilinear_repeat_clamp_96x80]",
    f"{P}::test_golden[ploc-spheres_preset_160]", f"{P}::test_golden[ploc-zfight_96]", f"{P}::test_golden[ploc-andreas_160x120]",
    f"{P}::test_golden[ploc-dragon_standin_c3_160x90]", f"{P}::test_golden_exhaustive[tr_def_96]",
    f"{P}::test_eight_lanes_per_ray_give_the_same_frame[cornell_c1_256]", f"{P}::test_eight_lanes_per_ray_give_the_same_frame[dragon_standin_c3_160x90]",
    f"{P}::test_eight_lanes_per_ray_give_the_same_frame[cornell_preset_sphere_192]",
    f"{P}::test_paths_and_levels_give_the_same_frame[monkey_192]", f"{P}::test_paths_and_levels_give_the_same_frame[teapot_d3_128x72]",
    f"{P}::test_glossy_rays_match_the_port[4-0-size0]", f"{P}::test_intersect_matches_oracle", f"{P}::test_wide_tree_answers_like_the_binary_one",
    f"{P}::test_axis_parallel_and_in_plane_rays", f"{P}::test_queue_overflow_is_clean_and_retried", f"{P}::test_batched_frame_equals_single_batch",
    f"{P}::test_sharded_tiles_compose", f"{P}::test_gather_frames_store_only_rows_with_hits", f"{P}::test_c5_lattice_small",
    f"{Q}::test_postprocess_inside_the_frame",
]


@pytest.mark.gpu
def test_checked_build_counts_no_out_of_range_index(rtb, tmp_path):
    if rtb.checked_build():
        pytest.skip("this whole session already runs against the checked build (conftest.py ends it with the verdict)")
    assert os.path.exists(CHECKED_LIB), f"{CHECKED_LIB} is missing: __graft_entry__.build() makes it (make EXTRA=-DRT_CHECKED=1 OUT=librtb200_checked.so BUILD=build_checked)"
    out = tmp_path / "violations.json"
    env = dict(os.environ, RTB200_LIB=CHECKED_LIB, RTB200_VIOLATIONS_OUT=str(out))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider"] + SUBSET, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    verdict = json.loads(out.read_text())
    assert verdict["checked"] is True, "the subprocess did not load the checked build"
    assert set(verdict["violations"]) == set(rtb.CHECK_SITES) and not any(verdict["violations"].values()), verdict
    print(f"checked build: {len(SUBSET)} tests, violations {verdict['violations']}")
