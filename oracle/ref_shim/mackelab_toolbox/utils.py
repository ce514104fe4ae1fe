"""Stand-ins for the three names the reference imports from `mackelab_toolbox.utils`."""
import sys


def total_size_handler(*types):
    def deco(f):
        return f
    return deco


def total_size(obj, *a, **k):
    return sys.getsizeof(obj)


def GitSHA(*a, **k):
    return None
