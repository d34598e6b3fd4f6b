from .modeling_utils import ModelMixin  # noqa


class AutoencoderKL:  # name only (isinstance / annotations in the reference)
    pass
