"""CPU: include/gvs.h is a plain C header (extern "C", no C++ types) and the adapter compiles as C++17 against the shim."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gvs_h_compiles_as_c99(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "gvs.h"\nint main(void) { gvs_extractor_params p; gvs_default_params(&p); return sizeof(gvs_keypoint) == 28 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L" + os.path.join(ROOT, "gtsam_vslam_b200"), "-lgvs", "-Wl,-rpath," + os.path.join(ROOT, "gtsam_vslam_b200")])
    assert subprocess.run([str(exe)]).returncode == 0


def test_adapter_compiles_with_shim(tmp_path):
    obj = tmp_path / "a.o"
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-c", os.path.join(ROOT, "tests", "cpp", "test_adapter.cpp"), "-o", str(obj)])
