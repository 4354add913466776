def list_physical_devices(kind=None):
    return []


def list_logical_devices(kind=None):
    return []


def set_visible_devices(*a, **k):
    pass


def set_logical_device_configuration(*a, **k):
    pass


class LogicalDeviceConfiguration:
    def __init__(self, memory_limit=None):
        self.memory_limit = memory_limit
