def create_device_mesh(shape, *a, **k):
    return [object()]
