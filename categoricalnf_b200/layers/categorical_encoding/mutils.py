"""Factory and CLI flags of the categorical encodings
(reference layers/categorical_encoding/mutils.py:14-72)."""
from .linear_encoding import LinearCategoricalEncoding
from .variational_dequantization import VariationalDequantization
from .variational_encoding import VariationalCategoricalEncoding


def add_encoding_parameters(parser, postfix=""):
    add = parser.add_argument
    add("--encoding_dim" + postfix, type=int, default=4, help="Dimensionality of the embeddings.")
    add("--encoding_dequantization" + postfix, action="store_true",
        help="Use variational dequantization for encoding categorical data.")
    add("--encoding_variational" + postfix, action="store_true",
        help="Use the variational encoding (joint encoder distribution with learned decoder).")
    add("--encoding_num_flows" + postfix, type=int, default=0, help="Number of flows in the embedding layer.")
    add("--encoding_hidden_layers" + postfix, type=int, default=2, help="Hidden layers of the encoding flows' nets.")
    add("--encoding_hidden_size" + postfix, type=int, default=128, help="Hidden size of the encoding flows' nets.")
    add("--encoding_num_mixtures" + postfix, type=int, default=8,
        help="Number of mixtures in the encoding coupling layers (if applicable).")
    add("--encoding_use_decoder" + postfix, action="store_true",
        help="Use a decoder instead of inverting all class-conditional flows for the likelihood.")
    add("--encoding_dec_num_layers" + postfix, type=int, default=1, help="Hidden layers of the decoder.")
    add("--encoding_dec_hidden_size" + postfix, type=int, default=64, help="Hidden size of the decoder.")


def encoding_args_to_params(args, postfix=""):
    get = lambda name: getattr(args, name + postfix)
    return {
        "use_dequantization": get("encoding_dequantization"),
        "use_variational": get("encoding_variational"),
        "use_decoder": get("encoding_use_decoder"),
        "num_dimensions": get("encoding_dim"),
        "flow_config": {"num_flows": get("encoding_num_flows"), "hidden_layers": get("encoding_hidden_layers"),
                        "hidden_size": get("encoding_hidden_size")},
        "decoder_config": {"num_layers": get("encoding_dec_num_layers"), "hidden_size": get("encoding_dec_hidden_size")},
    }


def create_encoding(encoding_params, dataset_class, vocab=None, vocab_size=-1, category_prior=None):
    assert not (vocab is None and vocab_size <= 0), \
        "[!] ERROR: When creating the encoding, either a torchtext vocabulary or the vocabulary size needs to be passed."
    use_dequantization = encoding_params.pop("use_dequantization")
    use_variational = encoding_params.pop("use_variational")
    if use_dequantization and "model_func" not in encoding_params["flow_config"]:
        # as upstream (:56-59): the CLI never supplies a network for the dequantization flow
        print("[#] WARNING: For using variational dequantization as encoding scheme, a model function needs to be specified"
              " in the encoding parameters, key \"flow_config\" which was missing here. Will deactivate dequantization...")
        use_dequantization = False
    if use_dequantization:
        encoding_flow = VariationalDequantization
    elif use_variational:
        encoding_flow = VariationalCategoricalEncoding
    else:
        encoding_flow = LinearCategoricalEncoding
    return encoding_flow(dataset_class=dataset_class, vocab=vocab, vocab_size=vocab_size,
                         category_prior=category_prior, **encoding_params)
