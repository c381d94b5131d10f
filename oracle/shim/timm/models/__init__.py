from .registry import _REG, register_model


def create_model(model_name, pretrained=False, checkpoint_path='', **kwargs):
    # timm 0.4.12 create_model(): look the factory up in the registry; never download here.
    return _REG[model_name](pretrained=False, **kwargs)
