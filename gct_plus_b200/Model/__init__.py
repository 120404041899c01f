from .sublayers import Sampler
from .layers import EncoderLayer, DecoderLayer
from .modules import Embeddings, PositionalEncoding, Norm, get_clones
from .vaetf import Vaetf
from .cvaetf import Cvaetf
