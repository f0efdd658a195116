"""single-rust-b200: B200-native sparse count-matrix analytics path of SingleRust."""
__version__ = "0.1.0"
