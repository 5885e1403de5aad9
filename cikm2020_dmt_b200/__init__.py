"""dmt-b200: B200-native DMT forward/backward ranking path."""
__version__ = "0.1.0"
