"""TEST INFRASTRUCTURE — builds the runnable copy of the reference used to pin the oracle.

    python oracle/build_ref.py [--tree /tmp/mmtpsm_ref] [--force]

The reference (/root/reference, read-only) neither imports nor compiles as shipped (SURVEY.md §0
F2/F3).  This script
  1. copies `maskrcnn_benchmark/`, `pycoco/` and `configs/` into a scratch tree OUTSIDE the repo
     (default /tmp/mmtpsm_ref) and applies the one-line patches of SURVEY.md §8(c) — no reference
     source is ever written into the repository;
  2. compiles the reference's own CPU `_C` extension (csrc/vision.cpp + csrc/cpu/*.cpp, i.e.
     nms_cpu + ROIAlign_forward_cpu) with g++ against this image's torch, and `pycocotools._mask`
     (pycoco/maskApi.c + _mask.pyx) with Cython;
  3. drops the two binaries into `oracle/_ref/` (git-ignored, travels to the GPU box) and into the
     scratch tree, and writes a ~60-line yacs-compatible `yacs/config.py` into the scratch tree.

`tests/golden/make_golden.py` imports the scratch tree to generate the committed fixtures.  Nothing
in the product path, the `-m gpu` tests, `smoke()` or `bench.py` reads /root/reference or the scratch
tree at run time; only `oracle/_ref/*.so` (binaries) are used there, as the checker / CPU baseline.
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")


def _sub(path, pattern, repl, count=0, must=True, flags=0):
    with open(path, "r", encoding="utf-8", errors="surrogateescape") as f:
        src = f.read()
    new, n = re.subn(pattern, repl, src, count=count, flags=flags)
    if must and n == 0:
        raise RuntimeError("patch did not apply: %s :: %s" % (path, pattern))
    with open(path, "w", encoding="utf-8", errors="surrogateescape") as f:
        f.write(new)


YACS = r'''
"""Minimal yacs-compatible CfgNode (yacs is not installed in this image)."""
import ast
import copy
import yaml


class CfgNode(dict):
    IMMUTABLE = "__immutable__"

    def __init__(self, init_dict=None):
        super().__init__()
        self.__dict__[CfgNode.IMMUTABLE] = False
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__[CfgNode.IMMUTABLE]:
            raise AttributeError("Attempted to set %s on an immutable CfgNode" % name)
        self[name] = value

    def is_frozen(self):
        return self.__dict__[CfgNode.IMMUTABLE]

    def _set_immutable(self, flag):
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_immutable(flag)

    def freeze(self):
        self._set_immutable(True)

    def defrost(self):
        self._set_immutable(False)

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        new = CfgNode()
        for k, v in self.items():
            dict.__setitem__(new, k, copy.deepcopy(v, memo))
        new.__dict__[CfgNode.IMMUTABLE] = self.__dict__[CfgNode.IMMUTABLE]
        return new

    @staticmethod
    def _decode(v):
        if not isinstance(v, str):
            return v
        try:
            return ast.literal_eval(v)
        except (ValueError, SyntaxError):
            return v

    @staticmethod
    def _coerce(new, old, key):
        if old is None or type(new) == type(old):
            return new
        if isinstance(old, tuple) and isinstance(new, list):
            return tuple(new)
        if isinstance(old, list) and isinstance(new, tuple):
            return list(new)
        if isinstance(old, float) and isinstance(new, int):
            return float(new)
        if isinstance(old, str) or new is None:
            return new
        raise ValueError("type mismatch for %s: %r vs %r" % (key, type(new), type(old)))

    def _merge(self, other, path):
        for k, v in other.items():
            full = ".".join(path + [k])
            if k not in self:
                raise KeyError("Non-existent config key: %s" % full)
            if isinstance(v, dict):
                self[k]._merge(v, path + [k])
            else:
                dict.__setitem__(self, k, self._coerce(self._decode(v), self[k], full))

    def merge_from_file(self, path):
        with open(path, "r") as f:
            self._merge(yaml.safe_load(f) or {}, [])

    def merge_from_other_cfg(self, other):
        self._merge(other, [])

    def merge_from_list(self, lst):
        assert len(lst) % 2 == 0
        for full, v in zip(lst[0::2], lst[1::2]):
            node = self
            keys = full.split(".")
            for k in keys[:-1]:
                node = node[k]
            if keys[-1] not in node:
                raise KeyError("Non-existent config key: %s" % full)
            dict.__setitem__(node, keys[-1], self._coerce(self._decode(v), node[keys[-1]], full))
'''


def make_tree(tree):
    if os.path.exists(tree):
        shutil.rmtree(tree)
    os.makedirs(tree)
    shutil.copytree(os.path.join(REF, "maskrcnn_benchmark"), os.path.join(tree, "maskrcnn_benchmark"))
    shutil.copytree(os.path.join(REF, "configs"), os.path.join(tree, "configs"))
    shutil.copytree(os.path.join(REF, "pycoco"), os.path.join(tree, "pycocotools"))
    subprocess.check_call(["chmod", "-R", "u+w", tree])
    os.makedirs(os.path.join(tree, "yacs"))
    open(os.path.join(tree, "yacs", "__init__.py"), "w").close()
    with open(os.path.join(tree, "yacs", "config.py"), "w") as f:
        f.write(YACS)
    open(os.path.join(tree, "pycocotools", "__init__.py"), "w").close()
    mb = os.path.join(tree, "maskrcnn_benchmark")
    p = lambda *a: os.path.join(mb, *a)
    # 1  torch >= 1.x dispatch API
    _sub(p("csrc/cpu/nms_cpu.cpp"), r"AT_DISPATCH_FLOATING_TYPES\(dets\.type\(\)", "AT_DISPATCH_FLOATING_TYPES(dets.scalar_type()")
    _sub(p("csrc/cpu/ROIAlign_cpu.cpp"), r"AT_DISPATCH_FLOATING_TYPES\(input\.type\(\)", "AT_DISPATCH_FLOATING_TYPES(input.scalar_type()")
    # .type().is_cuda() / .type()==.type() / .data<T>() still compile in torch 2.11 (deprecated)
    # 4/5  modules that are not in the tree
    _sub(p("modeling/roi_heads/roi_heads.py"), r"^(from .*maskiou_head.*)$", r"# \1", flags=re.M)
    _sub(p("modeling/relation/mask_relation_module.py"), r"^(from .*cuda_kmeans.*)$", r"# \1", flags=re.M)
    # 6  numpy >= 1.24
    _sub(p("modeling/rpn/anchor_generator.py"), r"np\.float\b", "float")
    # 7  torch._six
    _sub(p("utils/imports.py"), r"torch\._six\.PY3", "True")
    _sub(p("utils/c2_model_loading.py"), r"torch\._six\.PY3", "True", must=False)
    _sub(p("data/samplers/weighted_random_sampler.py"), r"^from torch\._six import int_classes as _int_classes", "_int_classes = int", must=False, flags=re.M)
    # 8  optional heavy imports that the path never uses
    _sub(p("utils/miscellaneous.py"), r"^from skimage import measure", "measure = None", must=False, flags=re.M)
    # 9  hard-coded cuda:0 in boxlist_nms
    _sub(p("structures/boxlist_ops.py"), r"boxes = boxes\.to\('cuda:0'\)", "pass")
    _sub(p("structures/boxlist_ops.py"), r"score = score\.to\('cuda:0'\)", "pass")
    # 12  .type(torch.cuda.FloatTensor)
    for f in ("modeling/relation/relation_module.py", "modeling/losses.py"):
        _sub(p(f), r"\.type\(torch\.cuda\.FloatTensor\)", ".float()", must=False)
    return mb


def patch_python_blocks(mb):
    """Patches that need more than a regex on one line (10, 11, 13 of SURVEY §8c)."""
    p = lambda *a: os.path.join(mb, *a)
    # 10  device shuffling in the feature extractors: make the `!= 0` test never fire on CPU tensors
    for f in ("modeling/roi_heads/box_head/roi_box_feature_extractors.py",
              "modeling/roi_heads/mask_head/roi_mask_feature_extractors.py"):
        _sub(p(f), r"device\s*=\s*x\[0\]\.device\.index", "device = x[0].device.index or 0")
    # 11  eval-mode mask head hands a list to the post-processor (F7)
    _sub(p("modeling/roi_heads/mask_head/mask_head.py"),
         r"result = self\.(post_processor|mask_generator)\(mask_logits,",
         r"result = self.\1(torch.cat(list(mask_logits)) if isinstance(mask_logits, (list, tuple)) else mask_logits,")
    # 13  CPU ROIAlign has no backward in the reference: route CPU tensors through the bit-identical
    #     torchvision op (aligned=False) so autograd works
    with open(p("layers/roi_align.py"), "a") as f:
        f.write('''

# ---- oracle patch (SURVEY §8c #13): CPU backward via torchvision (forward verified bit-identical)
import torchvision.ops as _tvops
_orig_roi_align_forward = ROIAlign.forward
def _cpu_roi_align_forward(self, input, rois):
    if input.is_cuda:
        return _orig_roi_align_forward(self, input, rois)
    return _tvops.roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio, aligned=False)
ROIAlign.forward = _cpu_roi_align_forward
''')


def build_C(mb, verbose=False):
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    csrc = os.path.join(mb, "csrc")
    srcs = [os.path.join(csrc, "vision.cpp")] + sorted(
        os.path.join(csrc, "cpu", f) for f in os.listdir(os.path.join(csrc, "cpu")) if f.endswith(".cpp"))
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(OUT, "_C" + ext)
    cmd = ["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-w", "-DTORCH_EXTENSION_NAME=_C",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           "-I", csrc, "-I", sysconfig.get_paths()["include"]]
    for inc in ce.include_paths():
        cmd += ["-isystem", inc]
    cmd += srcs + ["-o", out]
    for lp in ce.library_paths():
        cmd += ["-L", lp, "-Wl,-rpath," + lp]
    cmd += ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    shutil.copy(out, os.path.join(mb, "_C" + ext))
    return out


def build_pycoco(tree, verbose=False):
    import numpy as np
    pk = os.path.join(tree, "pycocotools")
    pyx = os.path.join(pk, "_mask.pyx")
    _sub(pyx, r"^# distutils: sources = .*$", "", flags=re.M)
    _sub(pyx, r"np\.NPY_OWNDATA", "np.NPY_ARRAY_OWNDATA", must=False)
    # nested comment in maskApi.c leaves a stray "*/" 
    with open(os.path.join(pk, "maskApi.c")) as f:
        lines = f.read().split("\n")
    for i in range(0, 12):
        if lines[i].strip() == "*/" and "/*" in lines[i - 1] and "*/" in lines[i - 1]:
            lines[i] = ""
    with open(os.path.join(pk, "maskApi.c"), "w") as f:
        f.write("\n".join(lines))
    subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", os.path.join(pk, "_mask.c")])
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(OUT, "_mask" + ext)
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-w", "-std=c99", "-I", pk, "-I", np.get_include(),
           "-I", sysconfig.get_paths()["include"], os.path.join(pk, "_mask.c"),
           os.path.join(pk, "maskApi.c"), "-o", out, "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    shutil.copy(out, os.path.join(pk, "_mask" + ext))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tree", default="/tmp/mmtpsm_ref")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    if not os.path.isdir(REF):
        print("reference not present (GPU box?) — using prebuilt oracle/_ref binaries")
        return 0
    mb = make_tree(a.tree)
    patch_python_blocks(mb)
    print("built", build_C(mb, a.verbose))
    print("built", build_pycoco(a.tree, a.verbose))
    print("reference tree:", a.tree)
    return 0


if __name__ == "__main__":
    sys.exit(main())
