"""Build the UNMODIFIED reference operators into oracle/_ref -- TEST INFRASTRUCTURE ONLY.

Recipe (no reference build system is run; no reference source enters the repo):
  * nvcc/g++ compile the reference's own native sources *where they lie* under
    /root/reference into two torch extension modules:
        oracle/_ref/pointnet2/_ext.so                      (pointnet2/_ext_src/src/*.{cpp,cu})
        oracle/_ref/pcdet/ops/iou3d_nms/iou3d_nms_cuda.so  (OpenPCDet/pcdet/ops/iou3d_nms/src/*)
  * the reference's Python operator modules are *installed* (copied, like
    `pip install --target` would) beside them so the stock code path can be run on
    the GPU box, where /root/reference does not exist; the reference's APPLICATION
    python (models/, utils/, dataset configs: the callers above the operator boundary)
    is installed into baseline/_ref/ (see install_python).
oracle/_ref/ and baseline/_ref/ are git-ignored (never enter history) but travel with gpurun.

Used by: tests (cross-check oracle == reference CUDA on the GPU box; reference CPU
BEV IoU golden vectors here), bench.py --impl reference.  Never by the product.

  python oracle/build_ref.py            # build + install
  python oracle/build_ref.py --dump-fma # print the fma contraction of the reference kernels
"""
import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(HERE, "_build", "ref_obj")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CXX = "/usr/bin/g++"


def _torch_flags():
    import torch
    from torch.utils.cpp_extension import include_paths, library_paths

    inc = include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
    lib = library_paths() + ["/usr/local/cuda/lib64"]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    return inc, lib, abi


def _compile(src, obj, name, inc, abi, extra_inc=()):
    common = ["-DTORCH_EXTENSION_NAME=%s" % name, "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=%d" % abi]
    incs = []
    for i in list(extra_inc) + inc:
        incs += ["-isystem" if i not in extra_inc else "-I", i]
    if src.endswith(".cu"):
        cmd = ["nvcc", "-ccbin", CXX, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
               "-w"] + ARCH + common + incs + ["-c", src, "-o", obj]
    else:
        cmd = [CXX, "-O2", "-std=c++17", "-fPIC", "-w"] + common + incs + ["-c", src, "-o", obj]
    subprocess.check_call(cmd)
    return obj


def _link(objs, so, lib):
    os.makedirs(os.path.dirname(so), exist_ok=True)
    cmd = [CXX, "-shared", "-o", so] + objs
    for l in lib:
        cmd += ["-L" + l, "-Wl,-rpath," + l]
    cmd += ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    subprocess.check_call(cmd)


def build_ext(name, src_dir, so, extra_inc=()):
    inc, lib, abi = _torch_flags()
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(f for f in os.listdir(src_dir) if f.endswith((".cpp", ".cu")))
    jobs = []
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        for f in srcs:
            obj = os.path.join(OBJ, "%s__%s.o" % (name, f.replace(".", "_")))
            jobs.append(ex.submit(_compile, os.path.join(src_dir, f), obj, name, inc, abi, extra_inc))
        objs = [j.result() for j in jobs]
    _link(objs, so, lib)
    return so


APP_OUT = os.path.join(os.path.dirname(HERE), "baseline", "_ref")


def install_python():
    """Install (copy, like `pip install --target` would) the reference's python files.

    oracle/_ref/   the reference OPERATOR stack: its python operator modules beside the two extensions built above
                   (the checker of the GPU tests, and the stack under `bench.py --impl reference`);
    baseline/_ref/ the reference APPLICATION: models/, utils/, the two dataset configs -- pure python ABOVE the operator
                   boundary (votenet_iou_branch.py, the loss helpers ...).  Both bench arms and the integration tests run
                   these unmodified callers, once on the drop-in operators and once on oracle/_ref.
    Both directories are git-ignored (no reference source enters the history) and travel with gpurun."""
    def cp(rel_src, dst_root, rel_dst=None):
        dst = os.path.join(dst_root, rel_dst or rel_src)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel_src), dst)

    for f in ("pointnet2_utils.py", "pointnet2_modules.py", "pytorch_utils.py"):
        cp("pointnet2/" + f, OUT)
    open(os.path.join(OUT, "pointnet2", "__init__.py"), "a").close()
    cp("OpenPCDet/pcdet/ops/iou3d_nms/iou3d_nms_utils.py", OUT, "pcdet/ops/iou3d_nms/iou3d_nms_utils.py")
    cp("OpenPCDet/pcdet/utils/common_utils.py", OUT, "pcdet/utils/common_utils.py")
    for d in ("pcdet", "pcdet/ops", "pcdet/ops/iou3d_nms", "pcdet/utils"):
        open(os.path.join(OUT, d, "__init__.py"), "a").close()
    for stale in ("models", "utils"):  # round-1 layout kept the callers inside oracle/_ref
        shutil.rmtree(os.path.join(OUT, stale), ignore_errors=True)
    for d in ("models", "utils"):
        for f in sorted(os.listdir(os.path.join(REF, d))):
            if f.endswith(".py"):
                cp(d + "/" + f, APP_OUT)
    cp("scannet/model_util_scannet.py", APP_OUT)
    cp("scannet/meta_data/scannet_means.npz", APP_OUT)
    cp("sunrgbd/model_util_sunrgbd.py", APP_OUT)
    cp("sunrgbd/sunrgbd_utils.py", APP_OUT)


def dump_fma():
    """Show how nvcc contracts the squared-distance / interpolation expressions of the reference."""
    inc, _, abi = _torch_flags()
    src_dir = os.path.join(REF, "pointnet2/_ext_src/src")
    os.makedirs(OBJ, exist_ok=True)
    for f in ("ball_query_gpu.cu", "sampling_gpu.cu", "interpolate_gpu.cu"):
        ptx = os.path.join(OBJ, f + ".ptx")
        incs = []
        for i in inc:
            incs += ["-isystem", i]
        subprocess.check_call(["nvcc", "-ccbin", CXX, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-w",
                               "-D_GLIBCXX_USE_CXX11_ABI=%d" % abi] + ARCH + incs +
                              ["-ptx", os.path.join(src_dir, f), "-o", ptx])
        print("==", f)
        for line in open(ptx):
            s = line.strip()
            if s.startswith((".visible .entry", "sub.f32", "mul.f32", "fma.rn.f32", "add.f32", "setp", "min.f32",
                             "cvt.f64.f32")):
                print("   ", s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dump-fma", action="store_true")
    ap.add_argument("--python-only", action="store_true", help="(re)install the python files, keep the built extensions")
    a = ap.parse_args()
    if not os.path.isdir(REF):
        print("reference tree %s not present; using prebuilt oracle/_ref if any" % REF)
        return 0
    if a.dump_fma:
        dump_fma()
        return 0
    if a.python_only:
        install_python()
        return 0
    src1 = os.path.join(REF, "pointnet2/_ext_src/src")
    so1 = build_ext("_ext", src1, os.path.join(OUT, "pointnet2", "_ext.so"),
                    extra_inc=[os.path.join(REF, "pointnet2/_ext_src/include")])
    src2 = os.path.join(REF, "OpenPCDet/pcdet/ops/iou3d_nms/src")
    so2 = build_ext("iou3d_nms_cuda", src2, os.path.join(OUT, "pcdet/ops/iou3d_nms", "iou3d_nms_cuda.so"))
    install_python()
    print("built", so1)
    print("built", so2)
    return 0


if __name__ == "__main__":
    sys.exit(main())
