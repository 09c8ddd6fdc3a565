from spatialrgpt_b200.conversation import *  # noqa: F401,F403
from spatialrgpt_b200.conversation import Conversation, SeparatorStyle, conv_templates, default_conversation  # noqa: F401
