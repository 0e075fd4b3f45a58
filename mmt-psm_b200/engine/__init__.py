from .MTtrainer import MTtrainer, weight_sum_losses, reduce_loss_dict  # noqa: F401
