"""Drop-in categorical encodings (reference layers/categorical_encoding/)."""
from .decoder import DecoderLinear, create_decoder, create_embed_layer
from .linear_encoding import LinearCategoricalEncoding
from .variational_encoding import VariationalCategoricalEncoding
from .variational_dequantization import VariationalDequantization
from .mutils import add_encoding_parameters, create_encoding, encoding_args_to_params

__all__ = ["DecoderLinear", "create_decoder", "create_embed_layer", "LinearCategoricalEncoding",
           "VariationalCategoricalEncoding", "VariationalDequantization", "add_encoding_parameters", "create_encoding", "encoding_args_to_params"]
