"""data_loaders/dataset.py:19-94 of the reference: the image helpers sample/generate_image.py imports (the training dataset
class of that file is outside the generation path)."""
from ...clip_encoder import CLIP_MEAN, CLIP_STD, crop_square, mask2bbox  # noqa: F401


def _convert_image_to_rgb(image):
    return image.convert("RGB")


def _transform(n_px):
    from torchvision.transforms import CenterCrop, Compose, Normalize, ToTensor
    return Compose([CenterCrop(n_px), _convert_image_to_rgb, ToTensor(), Normalize(CLIP_MEAN, CLIP_STD)])


def _transform_rgb(n_px):
    from torchvision.transforms import Compose, Normalize, Resize, ToTensor
    return Compose([ToTensor(), Normalize(CLIP_MEAN, CLIP_STD), Resize((n_px, n_px))])
