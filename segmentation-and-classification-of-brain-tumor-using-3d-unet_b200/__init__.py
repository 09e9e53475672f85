"""B200-native hot path of the enhanced 3D U-Net."""
