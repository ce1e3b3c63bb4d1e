"""The blob tables of the reference's SQLite store (``magmap/io/sqlite.py``), so that GPU
detections land in a database the reference's GUI and verifier read: same file schema
(version 4), same column order, same replace-on-duplicate rule.  Only what carries blobs and
the ROIs / experiments they hang off is mirrored."""
from __future__ import annotations

import datetime
import os
import sqlite3
from typing import Optional, Sequence, Tuple

import numpy as np

from ..cv import detector

DB_NAME_BASE = "magmap"
DB_SUFFIX_TRUTH = "_truth.db"
DB_VERSION = 4
_COLS_BLOBS = "roi_id, z, y, x, radius, confirmed, truth, channel"


def _create_tables(cur) -> None:
    """Schema of ``_create_db`` (sqlite.py:33-98)."""
    cur.execute("CREATE TABLE about (version INTEGER PRIMARY KEY, date DATE)")
    cur.execute("CREATE TABLE experiments (id INTEGER PRIMARY KEY AUTOINCREMENT, "
                "name TEXT, date DATE)")
    cur.execute("CREATE TABLE rois (id INTEGER PRIMARY KEY AUTOINCREMENT, "
                "experiment_id INTEGER, series INTEGER, offset_x INTEGER, offset_y INTEGER, "
                "offset_z INTEGER, size_x INTEGER, size_y INTEGER, size_z INTEGER, "
                "UNIQUE (experiment_id, series, offset_x, offset_y, offset_z))")
    cur.execute("CREATE TABLE blobs (id INTEGER PRIMARY KEY AUTOINCREMENT, "
                "roi_id INTEGER, x INTEGER, y INTEGER, z INTEGER, radius REAL, "
                "confirmed INTEGER, truth INTEGER, channel INTEGER, "
                "UNIQUE (roi_id, x, y, z, truth, channel))")
    cur.execute("CREATE TABLE blob_matches (id INTEGER PRIMARY KEY AUTOINCREMENT, "
                "roi_id INTEGER, blob1 INTEGER, blob2 INTEGER, dist REAL, "
                "FOREIGN KEY (roi_id) REFERENCES rois (id) ON UPDATE CASCADE ON DELETE CASCADE, "
                "FOREIGN KEY (blob1) REFERENCES blobs (id) ON UPDATE CASCADE ON DELETE CASCADE, "
                "FOREIGN KEY (blob2) REFERENCES blobs (id) ON UPDATE CASCADE ON DELETE CASCADE)")


def create_db(path: str) -> Tuple[sqlite3.Connection, sqlite3.Cursor]:
    """A new, empty database at ``path`` (an existing file is moved aside first)."""
    if os.path.exists(path):
        os.replace(path, path + ".bak")
    conn = sqlite3.connect(path)
    conn.row_factory = sqlite3.Row
    cur = conn.cursor()
    _create_tables(cur)
    cur.execute("INSERT INTO about (version, date) VALUES (?, ?)",
                (DB_VERSION, datetime.datetime.now().isoformat(" ")))
    conn.commit()
    return conn, cur


def open_db(path: str) -> Tuple[sqlite3.Connection, sqlite3.Cursor]:
    conn = sqlite3.connect(path)
    conn.row_factory = sqlite3.Row
    return conn, conn.cursor()


def insert_experiment(conn, cur, name: str, date=None) -> int:
    """sqlite.py:196-212."""
    if date is None:
        date = datetime.datetime.now()
    cur.execute("INSERT INTO experiments (name, date) VALUES (?, ?)",
                (name, date.isoformat(" ") if hasattr(date, "isoformat") else date))
    conn.commit()
    return cur.lastrowid


def insert_roi(conn, cur, exp_id: int, series: Optional[int], offset: Sequence[int],
               size: Sequence[int]):
    """``offset`` and ``size`` in x, y, z; a duplicate ROI is replaced (sqlite.py:241-267)."""
    if series is None:
        series = 0
    cur.execute("INSERT OR REPLACE INTO rois (experiment_id, series, offset_x, offset_y, "
                "offset_z, size_x, size_y, size_z) VALUES (?, ?, ?, ?, ?, ?, ?, ?)",
                (exp_id, series, *[int(v) for v in offset], *[int(v) for v in size]))
    conn.commit()
    return cur.lastrowid, "ROI inserted with offset {} and size {}".format(offset, size)


def select_or_insert_roi(conn, cur, exp_id: int, series: Optional[int], offset, size):
    """The id of the ROI with this experiment, offset and size (and series, when one is
    given), inserting it when there is none; returns ``(id, message)`` (sqlite.py:270-300)."""
    where = {"experiment_id": exp_id}
    where.update(zip(("offset_x", "offset_y", "offset_z"), (int(v) for v in offset)))
    where.update(zip(("size_x", "size_y", "size_z"), (int(v) for v in size)))
    if series is not None:
        where["series"] = series
    cur.execute("SELECT * FROM rois WHERE " + " AND ".join(f"{k} = ?" for k in where),
                list(where.values()))
    found = cur.fetchone()
    if found is None or len(found) == 0:
        return insert_roi(conn, cur, exp_id, series, offset, size)
    return found[0], "Found ROI {}".format(found[0])


def insert_blobs(conn, cur, roi_id: int, blobs) -> int:
    """Insert blobs given as ``z, y, x, radius, confirmed, truth, channel`` rows (seven
    columns, e.g. ``Blobs.blob_for_db``) under ``roi_id``; a blob equal in ``(roi_id, x, y,
    z, truth, channel)`` replaces the stored one (sqlite.py:359-384).  Returns the number
    of confirmed blobs."""
    rows, confirmed = [], 0
    for blob in np.asarray(blobs):
        rows.append([roi_id, *[v.item() if hasattr(v, "item") else v for v in blob]])
        if detector.Blobs.get_blob_confirmed(blob) == 1:
            confirmed += 1
    cur.executemany("INSERT OR REPLACE INTO blobs ({}) VALUES ({})".format(
        _COLS_BLOBS, ", ".join("?" * len(_COLS_BLOBS.split(", ")))), rows)
    conn.commit()
    return confirmed


def delete_blobs(conn, cur, roi_id: int, blobs) -> int:
    """Delete the blobs matching ``roi_id``, coordinates and channel (sqlite.py:387-412)."""
    deleted = 0
    for blob in np.asarray(blobs):
        cur.execute("DELETE FROM blobs WHERE roi_id = ? AND z = ? AND y = ? AND x = ? "
                    "AND channel = ?",
                    [roi_id, *[float(v) for v in blob[:3]],
                     float(detector.Blobs.get_blobs_channel(blob))])
        if cur.rowcount > 0:
            deleted += cur.rowcount
    conn.commit()
    return deleted


_BLOB_FIELDS = ("z", "y", "x", "radius", "confirmed", "truth", "channel")


def _parse_blobs(rows):
    """Rows -> ``(n, 7)`` table ``z, y, x, radius, confirmed, truth, channel`` and the row
    ids of the rows that carry one (sqlite.py:415-435)."""
    table = np.empty((len(rows), len(_BLOB_FIELDS)))
    for out, row in zip(table, rows):
        out[:] = [row[f] for f in _BLOB_FIELDS]
    ids = [row["id"] for row in rows if "id" in row.keys()]
    return table, ids


def select_blobs_by_roi(cur, roi_id: int):
    """sqlite.py:823-836."""
    cur.execute("SELECT * FROM blobs WHERE roi_id = ?", (roi_id,))
    return _parse_blobs(cur.fetchall())


def select_blobs_confirmed(cur, confirmed: int) -> np.ndarray:
    """sqlite.py:438-451."""
    cur.execute("SELECT {} FROM blobs WHERE confirmed = ?".format(_COLS_BLOBS), (confirmed,))
    return _parse_blobs(cur.fetchall())[0]
