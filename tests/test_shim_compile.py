"""CPU tier: immesh_b200/csrc/immesh_shim.hpp (the reference-signature forwarders: buildVoxelMap / updateVoxelMap / BuildResidualListOMP,
incremental_mesh_reconstruction, KD_TREE::Nearest_Search) goes through a compiler and the linker: compiled against minimal stand-ins
of the reference headers it needs (tests/shim_stubs: Eigen / PCL / voxel_loc.hpp members only) and linked with libimmesh_b200.so, so
every forwarded entry point exists with the argument types the shim passes.  The call sequence itself runs on the GPU tier through the
same C-ABI entry points (tests/test_parity_gpu.py::test_pv_entry_points, tests/test_mesh_gpu.py)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_compiles_and_links(tmp_path):
    from immesh_b200 import build
    lib = build.build_cuda()
    stubs = os.path.join(ROOT, "tests", "shim_stubs")
    exe = str(tmp_path / "shim_user")
    cmd = ["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror=return-type", "-I", stubs, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "immesh_b200", "csrc"),
           os.path.join(stubs, "shim_user.cpp"), "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib), "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True)     # main() only constructs the containers: no CUDA call
    assert r.returncode == 0, r.stderr[-1000:]
