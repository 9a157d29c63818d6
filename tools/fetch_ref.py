"""Copy the reference's shipped CHECKPOINTS (weights + their model_parameters.yml; no source code) into the git-ignored
``baseline/_ref/workdir`` so that they travel to the GPU box with the repo snapshot (gpurun ships git-ignored files).

The pretrained-regime parity tests and bench.py's ``--weights checkpoint`` read them from there; ``/root/reference`` does not
exist on the GPU box.  Run in the build container:  python tools/fetch_ref.py   (also called by __graft_entry__.build()).
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(os.environ.get('DDK_REFERENCE', '/root/reference'), 'workdir')
DST = os.path.join(ROOT, 'baseline', '_ref', 'workdir')
FILES = {
    'diffdockS_score_model': ('best_ema_inference_epoch_model.pt', 'model_parameters.yml'),
    'disco_diffdockS_score_model': ('best_ema_inference_epoch_model.pt', 'model_parameters.yml'),
    'disco_diffdockS_ar_model': ('best_model_loss.pt', 'model_parameters.yml'),
}


def fetch(verbose=True):
    if not os.path.isdir(SRC):
        return False
    for d, names in FILES.items():
        os.makedirs(os.path.join(DST, d), exist_ok=True)
        for n in names:
            s, t = os.path.join(SRC, d, n), os.path.join(DST, d, n)
            if os.path.exists(s) and (not os.path.exists(t) or os.path.getsize(t) != os.path.getsize(s)):
                shutil.copyfile(s, t)
                if verbose:
                    print('copied', os.path.relpath(t, ROOT))
    return True


if __name__ == '__main__':
    sys.exit(0 if fetch() else 1)
