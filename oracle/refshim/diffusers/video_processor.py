from .image_processor import VaeImageProcessor


class VideoProcessor(VaeImageProcessor):
    pass
