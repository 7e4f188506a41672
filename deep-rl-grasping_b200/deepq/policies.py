class MlpPolicy:
    unsupported = "DQN is outside the hot-path scope (DESIGN.md section 7)"
