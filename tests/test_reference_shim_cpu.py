"""CPU, build container only (skipped where /root/reference does not exist): the reference's OWN trainer code builds the
B200 networks after `deepards_b200.install(tad)`.

`deepards/train_ards_detector.py` is imported unmodified (its third-party imports that are absent from this image --
ventmap, imblearn, matplotlib, ... -- are satisfied by empty stub modules; none of them is touched by the methods used
here) and its unbound methods `BaseTraining.get_base_network` (:380-414), `<Model>.get_network` (:938-939, :963-964),
`BaseTraining.get_model` (:467-477) and `get_optimizer` (:416-422) run against a minimal stand-in for the trainer object.
Runs in a subprocess so that the stub modules never leak into the other tests."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SCRIPT = r'''
import sys, types, importlib.abc, importlib.machinery, argparse, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, %(ref)r); sys.path.insert(0, %(root)r)
STUBS = ("cv2", "matplotlib", "ventmap", "imblearn", "algorithms", "mock", "prettytable", "seaborn", "skimage")
class _Stub(types.ModuleType):
    __path__ = []
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})
class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in STUBS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
    def create_module(self, spec):
        return _Stub(spec.name)
    def exec_module(self, m):
        pass
sys.meta_path.insert(0, _Finder())
import torch
import deepards.train_ards_detector as tad          # the reference's trainer, unmodified
import deepards_b200 as D
ref_cnn_linear = tad.CNNLinearNetwork
D.install(tad)
assert tad.CNNLinearNetwork is D.CNNLinearNetwork and tad.CNNLinearNetwork is not ref_cnn_linear
assert "cnn_linear" in tad.network_map and "cnn_single_breath_linear" in tad.network_map

class Trainer(object):                                # what the methods below read from `self`
    n_metadata_inputs = 0
    is_2d_dataset = False
    is_2x1d_dataset = False
    model_cuda_wrapper = staticmethod(lambda m: m)    # no GPU in this container
for base, head_cls, model_cls in (("resnet18", D.CNNLinearNetwork, tad.CNNLinearModel),
                                  ("densenet18", D.CNNLinearNetwork, tad.CNNLinearModel),
                                  ("resnet18", D.CNNSingleBreathLinearNetwork, tad.CNNSingleBreathLinearModel)):
    t = Trainer()
    t.args = argparse.Namespace(base_network=base, load_base_network=None, initial_planes=64, resnet_first_pool_type="max",
                                resnet_double_conv=False, with_fft=False, only_fft=False, fft_real_only=False,
                                freeze_base_network=False, n_sub_batches=20, load_checkpoint=None, clip_grad=True,
                                clip_val=0.01, optimizer="sgd", learning_rate=1e-3, weight_decay=1e-4)
    t.get_base_network = lambda t=t: tad.BaseTraining.get_base_network(t)
    t.get_network = lambda bb, t=t, model_cls=model_cls: model_cls.get_network(t, bb)
    model = tad.BaseTraining.get_model(t)              # get_base_network -> get_network -> clamp hooks (:467-477)
    assert type(model) is head_cls, type(model)
    assert type(model.breath_block).__module__.startswith("deepards_b200."), type(model.breath_block)
    assert model.breath_block.network_name == base
    hooked = [p for p in model.parameters() if p._backward_hooks]
    assert len(hooked) == len(list(model.parameters()))          # the reference's clamp hook sits on every parameter
    opt = tad.BaseTraining.get_optimizer(t, model)                # torch.optim.SGD(momentum 0.9, nesterov) (:416-422)
    assert isinstance(opt, torch.optim.SGD) and opt.defaults["nesterov"] and opt.defaults["momentum"] == 0.9
    assert sum(p.numel() for g in opt.param_groups for p in g["params"]) == sum(p.numel() for p in model.parameters())
    try:
        model(torch.zeros(2, 20, 1, 224), None)
        raise SystemExit("a CPU forward must not succeed")
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
print("SHIM-OK")
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "deepards")), reason="the reference tree is only mounted in the build container")
def test_reference_trainer_builds_b200_networks_after_install():
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"ref": REF, "root": ROOT}], capture_output=True, text=True,
                       timeout=300, cwd="/tmp")
    assert r.returncode == 0 and "SHIM-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
