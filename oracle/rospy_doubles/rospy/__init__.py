"""TEST INFRASTRUCTURE: a header-only-style test double of `rospy`, just wide enough to import the
reference's local_navigator.py UNMODIFIED from /root/reference and drive its main loop
(oracle/ref_navigator.py).  Not part of the product; nothing under autonomouscar_b200/ imports it.

The double has no threads: messages are delivered from inside Rate.sleep(), which is one of the
interleavings a live rospy (callbacks on other threads) can produce.  The driver installs
`script` (what to deliver before each loop pass) and reads `published`.
"""


class ROSInterruptException(Exception):
    pass


class _State:
    def __init__(self):
        self.reset()

    def reset(self):
        self.subscribers = {}     # topic -> callback
        self.published = []       # (pass index, topic, dict of the message's fields)
        self.script = []          # list of lists of (topic, message); entry k is delivered in the k-th sleep
        self.passes = 0           # loop passes finished (= Rate.sleep() calls)


state = _State()


def loginfo(*_a, **_k):
    pass


logwarn = logerr = logdebug = loginfo


def get_caller_id():
    return "/local_navigation"


def init_node(*_a, **_k):
    pass


class Time:
    """rospy.Time.now(): the double's clock is the number of loop passes so far."""

    def __init__(self, secs=0.0):
        self.secs = secs

    @staticmethod
    def now():
        return Time(float(state.passes))


def is_shutdown():
    return state.passes > len(state.script)


class Subscriber:
    def __init__(self, topic, _type, callback, queue_size=None):
        state.subscribers[topic.lstrip("/")] = callback


class Publisher:
    def __init__(self, topic, _type, queue_size=None):
        self.topic = topic

    def publish(self, msg):
        # the reference re-uses ONE message object: snapshot its fields
        state.published.append((state.passes, self.topic, dict(vars(msg))))


class Rate:
    def __init__(self, hz):
        self.hz = hz

    def sleep(self):
        k = state.passes
        state.passes += 1
        if k < len(state.script):
            for topic, msg in state.script[k]:
                state.subscribers[topic.lstrip("/")](msg)
