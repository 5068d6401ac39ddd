"""Dictionary-size rule shared by both SAE variants.

The reference resolves the number of latents in one place (src/utils/models.py) and both model constructors call
it (l1autoencoder.py:48-50, topkautoencoder.py:52-54): a non-zero `n_dict_components` is taken literally, zero means
"expansion_factor times the activation size".  The name and the positional signature are part of the drop-in
surface; the bounds check is ours -- a negative size would otherwise surface much later as
an allocation error inside a kernel launch.
"""


def get_n_dict_components(activation_size: int, expansion_factor: int, n_dict_components: int) -> int:
    explicit = int(n_dict_components)
    if explicit < 0 or (explicit == 0 and (activation_size <= 0 or expansion_factor <= 0)):
        raise ValueError(f"cannot size a dictionary from activation_size={activation_size}, "
                         f"expansion_factor={expansion_factor}, n_dict_components={n_dict_components}")
    return explicit if explicit else activation_size * expansion_factor
