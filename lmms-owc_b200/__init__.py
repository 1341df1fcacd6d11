"""owc_b200 - B200-native scoring path (package body lives in lmms-owc_b200/)."""
