from .resize import interpolate_pos_embeddings, tf_bicubic_resize  # noqa: F401
