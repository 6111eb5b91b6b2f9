"""Import the UNMODIFIED ganecdotes reference from /root/reference on CPU.

Only used by `make_golden.py` (in the build container, where /root/reference is
mounted).  Nothing under tests/ that runs on the GPU box imports this module:
/root/reference does not exist there.

The reference's hot-path modules star-import host-only packages that are not
installed here (pylab/matplotlib, astropy, skimage, imageio, apex).  None of them
does arithmetic on the clustering path, so they are replaced by inert stub
modules.  The single exception is `apex.parallel.LARC`, whose arithmetic is on
the path: it is restated here from apex's published algorithm
(NVIDIA/apex `apex/parallel/LARC.py`, version unpinned by the reference) so the
reference's own `pretrain()` loop can run unmodified.
"""
import logging
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = "/root/reference"


class LARC(object):
    """Restatement of apex.parallel.LARC (clip=False path used by the reference,
    hfc_with_swav/swav_clustering.py:290-292)."""

    def __init__(self, optimizer, trust_coefficient=0.02, clip=True, eps=1e-8):
        self.optim = optimizer
        self.trust_coefficient = trust_coefficient
        self.eps = eps
        self.clip = clip

    def __getstate__(self):
        return self.optim.__getstate__()

    def __setstate__(self, state):
        self.optim.__setstate__(state)

    @property
    def state(self):
        return self.optim.state

    @property
    def param_groups(self):
        return self.optim.param_groups

    @param_groups.setter
    def param_groups(self, value):
        self.optim.param_groups = value

    def zero_grad(self):
        self.optim.zero_grad()

    def step(self):
        with torch.no_grad():
            weight_decays = []
            for group in self.optim.param_groups:
                weight_decay = group['weight_decay'] if 'weight_decay' in group else 0
                weight_decays.append(weight_decay)
                group['weight_decay'] = 0
                for p in group['params']:
                    if p.grad is None:
                        continue
                    param_norm = torch.norm(p.data)
                    grad_norm = torch.norm(p.grad.data)
                    if param_norm != 0 and grad_norm != 0:
                        adaptive_lr = self.trust_coefficient * param_norm / (
                            grad_norm + param_norm * weight_decay + self.eps)
                        if self.clip:
                            adaptive_lr = min(adaptive_lr / group['lr'], 1)
                        p.grad.data += weight_decay * p.data
                        p.grad.data *= adaptive_lr
        self.optim.step()
        for i, group in enumerate(self.optim.param_groups):
            group['weight_decay'] = weight_decays[i]


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    if "pylab" in sys.modules and getattr(sys.modules["pylab"], "_gx_stub", False):
        return
    plt = _stub("matplotlib.pyplot")
    mpl = _stub("matplotlib", pyplot=plt)
    _stub("matplotlib.patches", Ellipse=object, Rectangle=object)
    mpl.patches = sys.modules["matplotlib.patches"]
    # the reference relies on pylab's star-export for np / logging / plt
    # (lib/util/util.py:1,67)
    pylab = _stub("pylab", np=np, plt=plt, logging=logging, _gx_stub=True)
    for k in dir(np):
        if not k.startswith("_"):
            setattr(pylab, k, getattr(np, k))
    pylab.random = np.random
    _stub("astropy")
    _stub("astropy.io", fits=types.ModuleType("fits"))
    sys.modules["astropy.io.fits"] = sys.modules["astropy.io"].fits
    _stub("skimage")
    _stub("skimage.measure", regionprops=None)
    _stub("skimage.transform", rescale=None, resize=None)
    _stub("imageio")
    _stub("apex")
    _stub("apex.parallel")
    _stub("apex.parallel.LARC", LARC=LARC)
    try:
        import torch.utils.tensorboard  # noqa: F401
    except Exception:
        tb = _stub("torch.utils.tensorboard", SummaryWriter=object)
        torch.utils.tensorboard = tb


def load_reference():
    """Returns the reference modules needed on the clustering path."""
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import models.stylegan2.model as sg2_model
    import hfc_with_swav.swav_clustering as swav
    import lib.oneshot.image_augmentor as augm
    return types.SimpleNamespace(model=sg2_model, swav=swav, augmentor=augm)
