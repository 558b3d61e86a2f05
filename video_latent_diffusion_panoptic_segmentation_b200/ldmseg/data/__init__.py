from .synthetic import (block_majority_labels, ids_digest, split_cat_ins, teacher_ground_truth,  # noqa: F401
                        trained_like_rgb_latents)
