"""CPU: the drop-in shim (dropin/sitecustomize.py) makes the UNMODIFIED reference import this implementation
(SURVEY.md 8b "import paths").  Needs the reference tree, which only exists in the build container."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = "/root/reference/Our_UNet"


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not on this host")
def test_reference_trainer_imports_the_b200_model_through_the_shim():
    code = (
        "import models.unet as mu, models.losses as ml\n"
        "assert mu.UNet.__module__ == 'unet_implementations_b200.models.unet', mu.UNet.__module__\n"
        "assert ml.SimpleLoss.__module__ == 'unet_implementations_b200.models.losses'\n"
        "import src.train as T\n"                       # the reference's trainer, unmodified (train.py:28-29)
        "assert T.UNet is mu.UNet and T.SimpleLoss is ml.SimpleLoss\n"
        "assert T.create_optimizer.__module__ == 'src.train'\n"
        "import src.models.unet as smu\n"               # evaluate.py:32
        "assert smu.UNet is mu.UNet\n"
        "m = T.UNet()\n"
        "opt = T.create_optimizer(m, 0.01, 1e-4, 0.99)\n"
        "assert len(opt.param_groups[0]['params']) == len(m.state_dict()) == 90\n"
        "print('ok')\n"
    )
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT, REF])
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp", timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-4000:]
