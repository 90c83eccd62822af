"""Import the upstream EgoNet modules from /root/reference (build container only).

Used exclusively by ``make_golden.py`` to generate the committed fixtures.
matplotlib is absent from the image and only used for plotting upstream, so
empty stand-in modules are injected.
"""
import sys
import types

REF_ROOT = '/root/reference'


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def import_reference():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    mpl = _stub('matplotlib')
    plt = _stub('matplotlib.pyplot', ion=lambda: None)
    mpl.pyplot = plt
    _stub('matplotlib.patches')
    _stub('mpl_toolkits')
    _stub('mpl_toolkits.mplot3d', Axes3D=object)
    import libs.model.heatmapModel.hrnet as hrnet
    import libs.model.FCmodel as fcmodel
    import libs.common.img_proc as img_proc
    import libs.common.transformation as transformation
    import libs.dataset.normalization.operations as operations
    import libs.model.egonet as egonet
    return dict(hrnet=hrnet, fcmodel=fcmodel, img_proc=img_proc,
                transformation=transformation, operations=operations, egonet=egonet)
