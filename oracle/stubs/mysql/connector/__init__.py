"""Stand-in for mysql.connector: connect() always fails soft, as with no DB server."""


class Error(Exception):
    pass


def connect(**kw):
    raise Error("no database in the golden-vector harness")
