class AxesImage:
    pass
