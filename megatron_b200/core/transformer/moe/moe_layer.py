"""placeholder — replaced below in the same commit series"""
from ..module import MegatronModule


class BaseMoELayer(MegatronModule):
    pass


class MoELayer(BaseMoELayer):
    pass
