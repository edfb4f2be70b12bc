// writer.cpp -- see writer.hpp
#include <stdio.h>
#include <string.h>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include "writer.hpp"

struct MxWriter
{
   struct Lane { std::deque<MxWriteTask> q; std::thread th; };
   std::vector<Lane> lanes;
   std::mutex mu;
   std::condition_variable cv_work, cv_done;
   int pending[2] = {0, 0};
   int in_flight = 0;
   bool stop = false;
   bool failed = false;
   std::string error;
};

static bool run_task (const MxWriteTask &t, std::string &err)
{
   FILE *fp = fopen (t.path.c_str (), t.create ? "w+b" : "r+b");
   if (fp == nullptr) { err = std::string (t.create ? "unable to create " : "unable to open ") + t.path; return false; }
   bool ok = t.create ? (32 == fwrite (t.header, 1, 32, fp)) : (0 == fseek (fp, 0, SEEK_END));
   ok = ok && ((t.bytes == 0) || (t.bytes == fwrite (t.data, 1, t.bytes, fp)));
   ok = ok && (0 == fseek (fp, 20, SEEK_SET)) && (4 == fwrite (t.rows_be, 1, 4, fp));
   if ((0 != fclose (fp)) || !ok) { err = "write error on " + t.path; return false; }
   return true;
}

static void lane_main (MxWriter *w, int k)
{
   std::unique_lock<std::mutex> lk (w->mu);
   while (true)
     {
        w->cv_work.wait (lk, [&] { return w->stop || !w->lanes[k].q.empty (); });
        if (w->lanes[k].q.empty ()) { if (w->stop) return; continue; }
        MxWriteTask t = w->lanes[k].q.front ();
        w->lanes[k].q.pop_front ();
        lk.unlock ();
        std::string err;
        const bool ok = run_task (t, err);
        lk.lock ();
        if (!ok && !w->failed) { w->failed = true; w->error = err; }
        w->pending[t.buffer & 1] -= 1;
        w->in_flight -= 1;
        w->cv_done.notify_all ();
     }
}

MxWriter *mxw_create (int n_threads)
{
   if (n_threads < 1) n_threads = 1;
   if (n_threads > 32) n_threads = 32;
   MxWriter *w = new MxWriter ();
   w->lanes.resize ((size_t) n_threads);
   for (int k = 0; k < n_threads; k++) w->lanes[k].th = std::thread (lane_main, w, k);
   return w;
}
void mxw_submit (MxWriter *w, int lane, const MxWriteTask &t)
{
   std::lock_guard<std::mutex> lk (w->mu);
   w->lanes[(size_t) lane % w->lanes.size ()].q.push_back (t);
   w->pending[t.buffer & 1] += 1;
   w->in_flight += 1;
   w->cv_work.notify_all ();
}
void mxw_wait_buffer (MxWriter *w, int buffer)
{
   std::unique_lock<std::mutex> lk (w->mu);
   w->cv_done.wait (lk, [&] { return w->pending[buffer & 1] == 0; });
}
int mxw_flush (MxWriter *w, std::string *err)
{
   std::unique_lock<std::mutex> lk (w->mu);
   w->cv_done.wait (lk, [&] { return w->in_flight == 0; });
   if (w->failed) { if (err) *err = w->error; w->failed = false; return -1; }
   return 0;
}
int mxw_failed (MxWriter *w, std::string *err)
{
   std::lock_guard<std::mutex> lk (w->mu);
   if (w->failed) { if (err) *err = w->error; return 1; }
   return 0;
}
void mxw_destroy (MxWriter *w)
{
   if (w == nullptr) return;
   mxw_flush (w, nullptr);
   {
      std::lock_guard<std::mutex> lk (w->mu);
      w->stop = true;
      w->cv_work.notify_all ();
   }
   for (auto &l : w->lanes) if (l.th.joinable ()) l.th.join ();
   delete w;
}
