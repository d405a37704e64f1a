from .distance_descriptors import *
