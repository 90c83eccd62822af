"""``python tools/inference.py --cfg <yml> [--visualize B] [--batch_to_show N]``

Same entry point and control flow as the reference's ``tools/inference.py`` (``main`` :215-284,
``inference`` :135-199), with ``libs.model.*`` served by the B200-native mirror.  Everything that is
NOT on the per-crop hot path -- YAML/argument parsing, logging, the KITTI dataset and detector-box
reader, the C++ evaluator -- is used from a checkout of the reference, found through
``EGONET_REFERENCE`` (default ``../reference`` next to this repo or ``/root/reference``).
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def _reference_root():
    for cand in (os.environ.get('EGONET_REFERENCE'), os.path.join(os.path.dirname(ROOT), 'reference'),
                 '/root/reference'):
        if cand and os.path.isdir(os.path.join(cand, 'libs')):
            return cand
    raise SystemExit('set EGONET_REFERENCE to a checkout of Nicholasli1995/EgoNet (dataset, logger and '
                     'evaluator code is used from there; only libs.model.* runs natively)')


def install_native_model():
    """Alias the native mirror over the reference's model modules (INTEGRATION.md option A)."""
    for name in ('model', 'model.heatmapModel', 'model.heatmapModel.hrnet', 'model.FCmodel', 'model.egonet'):
        sys.modules['libs.' + name] = importlib.import_module('egonet_b200.libs.' + name)


def main():
    ref = _reference_root()
    sys.path.insert(0, ref)
    import libs  # noqa: F401  (the reference package: arguments, logger, dataset ...)
    install_native_model()
    import torch
    if not torch.cuda.is_available():
        raise ValueError('CPU-based inference is not maintained.')       # same guard as upstream :227-231
    sys.argv[0] = os.path.join(ref, 'tools', 'inference.py')
    os.chdir(os.path.join(ref, 'tools'))                                   # upstream uses paths relative to tools/
    import runpy
    runpy.run_path(sys.argv[0], run_name='__main__')


if __name__ == '__main__':
    main()
