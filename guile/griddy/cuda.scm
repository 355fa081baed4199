;;; (griddy cuda) — Guile side of the B200 actor-update path.
;;;
;;; Binds libgriddy_cuda.so (include/griddy_cuda.h) with (system foreign), flattens a griddy world
;;; into the structure-of-arrays bytevectors the C ABI takes, and rebuilds GOOPS actors from a
;;; read-back so that `draw' (griddy/simulate.scm:155-160) keeps working.
;;;
;;; NOT EXECUTED UNDER GUILE IN THIS REPOSITORY (the build image has none).  It IS executed by
;;; tests/test_guile_glue.py under the Scheme evaluator that records the golden fixtures, with (system
;;; foreign) replaced by a stand-in for the library: the arrays this file hands to griddy_upload_network /
;;; _geometry / _actors for the reference's own examples equal the fixtures' (which the device tests
;;; upload), and cuda-download-world! rebuilds the reference's world from a read-back.  The same symbols,
;;; argument order and array layouts are exercised through ctypes by griddy_b200/device.py.
;;;
;;; Conventions (include/griddy_cuda.h): item id = registration index = N-1-i for position i of
;;; (ref world 'static-items) (world.scm:21-22 conses); actor ids such that linking the actors in id
;;; order rebuilds every segment's list: the k-th actor of (get-actors world) gets id M-1-k.
(define-module (griddy cuda)
  #:use-module (system foreign)
  #:use-module (rnrs bytevectors)
  #:use-module (srfi srfi-1)
  #:use-module (srfi srfi-69)
  #:use-module (oop goops)
  #:use-module (griddy constants)
  #:use-module (griddy util)
  #:use-module (griddy core)
  #:use-module (griddy simulate route)
  #:use-module (chickadee math vector)
  #:use-module (chickadee math bezier)
  #:export (cuda-open cuda-close cuda-upload-world! cuda-step! cuda-step-async! cuda-download-world!
            cuda-last-step-ms cuda-positions make-frame-buffer cuda-frame-begin! cuda-frame-end!
            make-delta-buffer cuda-delta-begin! cuda-delta-end! cuda-delta-apply!))

(define %lib (dynamic-link (or (getenv "GRIDDY_CUDA_LIB") "libgriddy_cuda")))

(define (c-fn name ret . args)
  (pointer->procedure ret (dynamic-func name %lib) args))

(define %create   (c-fn "griddy_create" int '* int))
(define %destroy  (c-fn "griddy_destroy" void '*))
(define %error    (c-fn "griddy_last_error" '* '*))
(define %const    (c-fn "griddy_set_constants" int '* double double))
(define %network  (c-fn "griddy_upload_network" int '* int64 '* '* '* '* '* '* '* '*))
(define %actors   (c-fn "griddy_upload_actors" int '* int64 '* '* '* '* '* '* '* '* '* '* '* '* '* '*))
(define %step     (c-fn "griddy_step" int '* int64))
(define %step-async (c-fn "griddy_step_async" int '* int64))
(define %route-graph (c-fn "griddy_set_route_graph" int '* '* '* '* '*))
(define %download (c-fn "griddy_download_actors" int '* '* '* '* '* '* '* '* '* '* '* '*))
(define %step-ms  (c-fn "griddy_last_step_ms" double '*))
(define %geometry (c-fn "griddy_upload_geometry" int '* '*))
(define %positions (c-fn "griddy_download_positions" int '* int64 '* '* '* '*))
(define %frame-begin (c-fn "griddy_positions_begin" int '* '* int64))
(define %frame-end (c-fn "griddy_positions_end" int '* '*))
(define %delta-bytes (c-fn "griddy_delta_frame_bytes" int64 int64))
(define %delta-begin (c-fn "griddy_delta_begin" int '* '* int64))
(define %delta-end (c-fn "griddy_delta_end" int '* '* '*))
(define %delta-apply (c-fn "griddy_delta_apply" int '* '* int64))
(define %host-alloc (c-fn "griddy_host_alloc" '* int64))
(define %slots    (c-fn "griddy_slots_in_use" int64 '*))

(define (check ctx rc)
  ;; nonzero return -> (throw 'griddy-cuda code message), the reference's throw style (core.scm:83)
  (unless (zero? rc)
    (throw 'griddy-cuda rc (pointer->string (%error ctx)))))

(define (cuda-open)
  (let ((slot (make-bytevector (sizeof '*) 0)))
    (let ((rc (%create (bytevector->pointer slot) -1)))
      (unless (zero? rc) (throw 'griddy-cuda rc "griddy_create failed (no CUDA device?)")))
    (let ((ctx (make-pointer (bytevector-uint-ref slot 0 (native-endianness) (sizeof '*)))))
      (check ctx (%const ctx
                         (exact->inexact *simulate/time-step*)     ; constants.scm:38
                         (exact->inexact (* 2 *actor/size*))))     ; route.scm:62
      ctx)))

(define (cuda-close ctx) (%destroy ctx))
(define (cuda-last-step-ms ctx) (%step-ms ctx))

;;; ---- typed bytevector helpers ---------------------------------------------------------------
(define (make-u8  n) (make-bytevector n 0))
(define (make-i32 n fill) (let ((bv (make-bytevector (* 4 n))))
                            (do ((i 0 (1+ i))) ((= i n) bv) (bytevector-s32-native-set! bv (* 4 i) fill))))
(define (make-i64 n) (make-bytevector (* 8 n) 0))
(define (make-f64 n) (make-bytevector (* 8 n) 0))
(define (i32! bv i v) (bytevector-s32-native-set! bv (* 4 i) v))
(define (i64! bv i v) (bytevector-s64-native-set! bv (* 8 i) v))
(define (f64! bv i v) (bytevector-ieee-double-native-set! bv (* 8 i) (exact->inexact v)))
(define (ptr bv) (bytevector->pointer bv))
(define (dir->int d) (if (eq? d 'forw) 0 1))

;;; ---- flatten the static network -------------------------------------------------------------
(define (registration-index world)
  "eq-hash: static item -> registration index (world.scm:21-22)"
  (let* ((items (get-static-items world))
         (n (length items))
         (table (make-hash-table eq?)))
    (let loop ((l items) (i 0))
      (unless (null? l)
        (hash-table-set! table (car l) (- n 1 i))
        (loop (cdr l) (1+ i))))
    table))

(define (upload-network! ctx world index)
  (let* ((n (hash-table-size index))
         (kind (make-u8 n)) (len (make-f64 n)) (dir (make-u8 n))
         (seg (make-i32 n -1)) (out (make-i32 n -1)) (jun (make-i32 n -1))
         (of (make-i32 n -1)) (ob (make-i32 n -1)))
    (define (id x) (hash-table-ref index x))
    (define (outer segment side)
      (catch 'road-has-no-lanes
        (lambda () (id (ref (off-road->on-road
                             (make <location/off-road> #:road-segment segment
                                   #:road-side-direction side #:pos-param 0.0))
                            'road-lane)))
        (lambda _ -1)))
    (hash-table-walk
     index
     (lambda (item r)
       (cond
        ((is-a? item <road-junction>) (bytevector-u8-set! kind r 0))
        ((is-a? item <road-segment>)
         (bytevector-u8-set! kind r 1)
         (i32! of r (outer item 'forw))                ; location.scm:16-24
         (i32! ob r (outer item 'back)))
        ((is-a? item <road-lane/segment>)
         (bytevector-u8-set! kind r 2)
         (f64! len r (get-length item))                ; spatial.scm:28-32, host-side fp64
         (bytevector-u8-set! dir r (dir->int (ref item 'direction)))
         (i32! seg r (id (ref item 'segment))))
        ((is-a? item <road-lane/junction>)
         (bytevector-u8-set! kind r 3)
         (f64! len r (get-length item))                ; spatial.scm:34-46
         (i32! out r (id (get-segment-lane item)))     ; core.scm:106-111
         (i32! jun r (id (ref item 'junction))))
        (else (bytevector-u8-set! kind r 4)))))
    (check ctx (%network ctx n (ptr kind) (ptr len) (ptr dir) (ptr seg) (ptr out) (ptr jun)
                         (ptr of) (ptr ob)))))

;;; ---- geometry for the drawing read-back -----------------------------------------------------
;;; 8 doubles per item (include/griddy_cuda.h, griddy_upload_geometry), every number produced by the
;;; reference's own procedures, so that the device only evaluates origin + pos-param * vec (and
;;; bezier-curve-point-at on junction lanes) — spatial.scm:125-146.
(define (upload-geometry! ctx world index)
  (let* ((n (hash-table-size index))
         (geom (make-f64 (* 8 n))))
    (define (vec! r at v)
      (f64! geom (+ (* 8 r) at) (vec2-x v))
      (f64! geom (+ (* 8 r) at 1) (vec2-y v)))
    (define (off-road-origin segment side)
      (get-pos (make <location/off-road> #:road-segment segment
                     #:road-side-direction side #:pos-param 0)))
    (hash-table-walk
     index
     (lambda (item r)
       (cond
        ((is-a? item <road-lane/segment>)
         (vec! r 0 (get-pos item 'beg))                 ; spatial.scm:104-107
         (vec! r 2 (get-vec item)))                     ; spatial.scm:156-158
        ((is-a? item <road-lane/junction>)
         (let ((curve (ref item 'curve)))               ; core.scm:196-206
           (vec! r 0 (bezier-curve-p0 curve))
           (vec! r 2 (bezier-curve-p1 curve))
           (vec! r 4 (bezier-curve-p2 curve))
           (vec! r 6 (bezier-curve-p3 curve))))
        ((is-a? item <road-segment>)
         (vec! r 0 (off-road-origin item 'forw))        ; v-beg + v-offset, spatial.scm:125-136
         (vec! r 2 (get-vec item))
         (vec! r 4 (off-road-origin item 'back))))))
    (check ctx (%geometry ctx (ptr geom)))))

;;; ---- the lane graph find-route searches, for find-route on the device ------------------------
;;; (griddy_set_route_graph) neighbours, cost and midpoints exactly as route.scm:19-40 defines them, every number
;;; from the reference's own procedures; the device then runs the A* itself when the actors are uploaded
;;; without routes.
(define (upload-route-graph! ctx world index)
  (let* ((n (hash-table-size index))
         (items (list->vector (reverse (get-static-items world))))   ; id -> item
         (nbr-off (make-i64 (1+ n)))
         (cost (make-f64 n)) (mid (make-f64 (* 2 n)))
         (nbrs '()) (k 0))
    (define (id x) (hash-table-ref index x))
    (do ((r 0 (1+ r))) ((= r n))
      (let ((item (vector-ref items r)))
        (i64! nbr-off r k)
        (cond
         ((is-a? item <road-lane/segment>)
          (for-each (lambda (jl) (set! nbrs (cons (id jl) nbrs)) (set! k (1+ k)))
                    (get-junction-lanes item))                      ; route.scm:20-21, core.scm:97-104
          (f64! cost r (get-length (ref item 'segment))))            ; route.scm:26-29
         ((is-a? item <road-lane/junction>)
          (set! nbrs (cons (id (get-segment-lane item)) nbrs))       ; route.scm:22-23, core.scm:106-111
          (set! k (1+ k))))                                          ; cost 0, route.scm:31-34
        (when (is-a? item <road-lane>)
          (let ((m (get-midpoint item)))                             ; route.scm:37-40, spatial.scm:68-73
            (f64! mid (* 2 r) (vec2-x m))
            (f64! mid (1+ (* 2 r)) (vec2-y m))))))
    (i64! nbr-off n k)
    (let ((nbr (make-i32 (max 1 k) -1)))
      (let fill ((l (reverse nbrs)) (i 0))
        (unless (null? l) (i32! nbr i (car l)) (fill (cdr l) (1+ i))))
      (check ctx (%route-graph ctx (ptr nbr-off) (ptr nbr) (ptr cost) (ptr mid))))))

;;; ---- flatten actors, agendas and routes -----------------------------------------------------
(define* (upload-actors! ctx world index #:key (routes 'host))
  ;; id order is the reverse of get-actors order (see header)
  (let* ((actors (reverse (get-actors world)))
         (n (length actors))
         (n-items (apply + (map (lambda (a) (length (ref a 'agenda))) actors)))
         (loc-kind (make-u8 n)) (container (make-i32 n -1)) (pos (make-f64 n)) (side (make-u8 n))
         (speed (make-f64 n)) (speed-dt (make-f64 n)) (agenda-off (make-i64 (1+ n)))
         (item-kind (make-u8 n-items)) (item-sleep (make-i32 n-items 0))
         (item-seg (make-i32 n-items -1)) (item-side (make-u8 n-items)) (item-pos (make-f64 n-items))
         (route-off (make-i64 (1+ n-items)))
         (route-lanes '()) (n-route 0) (k 0))
    (define (id x) (hash-table-ref index x))
    (let loop ((l actors) (a 0))
      (unless (null? l)
        (let* ((actor (car l)) (loc (ref actor 'location)) (cur #f))
          ;; an actor in the middle of a route cannot be uploaded: the device starts every actor with route 'none
          ;; (actor.scm:19-21), which is what add-actors! leaves behind
          (unless (eq? (ref actor 'route) 'none)
            (throw 'griddy-cuda 1 "cuda-upload-world!: an actor is already on a route; upload the world add-actors! built"))
          (if (is-a? loc <location/on-road>)
              (begin (bytevector-u8-set! loc-kind a 1)
                     (i32! container a (id (ref loc 'road-lane))))
              (begin (i32! container a (id (ref loc 'road-segment)))
                     (bytevector-u8-set! side a (dir->int (ref loc 'road-side-direction)))
                     (set! cur loc)))
          (f64! pos a (ref loc 'pos-param))
          (f64! speed a (ref actor 'max-speed))
          ;; exact rational times exact integer stays exact, then one rounding (route.scm:64)
          (f64! speed-dt a (* (ref actor 'max-speed) *simulate/time-step*))
          (i64! agenda-off a k)
          (for-each
           (lambda (item)
             (case (car item)
               ((sleep-for) (i32! item-sleep k (cadr item)))
               ((travel-to)
                (let ((dest (cadr item)))
                  (bytevector-u8-set! item-kind k 1)
                  (i32! item-seg k (id (ref dest 'road-segment)))
                  (bytevector-u8-set! item-side k (dir->int (ref dest 'road-side-direction)))
                  (f64! item-pos k (ref dest 'pos-param))
                  ;; the reference's own find-route (route.scm:17-51), evaluated at upload: the start
                  ;; of travel item k is where the previous travel ended (simulate.scm:87-91)
                  (when (and cur (eq? routes 'host))
                    (let ((steps (find-route (off-road->on-road cur) (off-road->on-road dest))))
                      (for-each (lambda (step)
                                  (when (eq? (car step) 'turn-onto)
                                    (set! route-lanes (cons (id (cadr step)) route-lanes))
                                    (set! n-route (1+ n-route))))
                                steps)))
                  (set! cur (on-road->off-road (off-road->on-road dest)))))
               (else (throw 'unknown-agenda-item item)))
             (set! k (1+ k))
             (i64! route-off k n-route))
           (ref actor 'agenda)))
        (loop (cdr l) (1+ a))))
    (i64! agenda-off n k)
    (let ((route-lane (make-i32 (max 1 n-route) -1)))
      (let fill ((l (reverse route-lanes)) (i 0))
        (unless (null? l) (i32! route-lane i (car l)) (fill (cdr l) (1+ i))))
      ;; routes 'device: no route lists cross the boundary, the device's find-route fills them in
      (check ctx (%actors ctx n (ptr loc-kind) (ptr container) (ptr pos) (ptr side) (ptr speed)
                          (ptr speed-dt) (ptr agenda-off) (ptr item-kind) (ptr item-sleep)
                          (ptr item-seg) (ptr item-side) (ptr item-pos)
                          (if (eq? routes 'host) (ptr route-off) %null-pointer)
                          (if (eq? routes 'host) (ptr route-lane) %null-pointer))))
    actors))

(define* (cuda-upload-world! ctx world #:key (routes 'host))
  "Flatten WORLD and upload it.  Returns (index . actors): what cuda-download-world! needs.
ROUTES: 'host — every travel-to's route comes from the reference's own find-route, evaluated here;
'device — the lane graph is uploaded and the device's find-route (the same A*) computes them."
  (let ((index (registration-index world)))
    (upload-network! ctx world index)
    (upload-geometry! ctx world index)
    (when (eq? routes 'device) (upload-route-graph! ctx world index))
    (cons index (upload-actors! ctx world index #:routes routes))))

(define (cuda-step! ctx n-ticks)
  (check ctx (%step ctx n-ticks)))

(define (cuda-step-async! ctx n-ticks)
  "Enqueue N-TICKS ticks and return without waiting for the device (frame loops); a device-side error is thrown by a
later call, at the latest by the next (cuda-step! ctx 0)."
  (check ctx (%step-async ctx n-ticks)))

;;; ---- read the world back into a fresh skeleton ----------------------------------------------
(define (cuda-download-world! ctx handle make-skeleton)
  "Build a new world from (make-skeleton) holding the device's actors, for `draw'."
  (let* ((actors (cdr handle))
         (n (length actors))
         (world++ (make-skeleton))
         (items (list->vector (reverse (get-static-items world++))))   ; id -> item
         (alive (make-u8 n)) (loc-kind (make-u8 n)) (container (make-i32 n -1)) (side (make-u8 n))
         (pos (make-f64 n)) (order (make-i32 n -1)) (n-order (make-i64 1)))
    (check ctx (%download ctx (ptr alive) (ptr loc-kind) (ptr container) (ptr side) (ptr pos)
                          %null-pointer %null-pointer %null-pointer %null-pointer
                          (ptr order) (ptr n-order)))
    ;; link! in reverse get-actors order so that every segment's list comes out as on the device
    (let ((actor-of (list->vector actors)))
      (do ((k (1- (bytevector-s64-native-ref n-order 0)) (1- k))) ((< k 0))
        (let* ((a (bytevector-s32-native-ref order (* 4 k)))
               (old (vector-ref actor-of a))
               (new (make <actor> #:max-speed (ref old 'max-speed)))
               (item (vector-ref items (bytevector-s32-native-ref container (* 4 a))))
               (p (bytevector-ieee-double-native-ref pos (* 8 a))))
          (link! new
                 (if (= 1 (bytevector-u8-ref loc-kind a))
                     (make <location/on-road> #:road-lane item #:pos-param p)
                     (make <location/off-road> #:road-segment item #:pos-param p
                           #:road-side-direction (if (zero? (bytevector-u8-ref side a)) 'forw 'back)))))))
    world++))

;;; ---- the drawing read-back: (get-pos actor) of every actor, computed on the device -----------
;;; `draw' needs nothing else of an actor (draw.scm:74-76).  One f32vector of x, y pairs per frame,
;;; in slot order; a pair of NaNs marks a slot without a living actor.
(define (cuda-positions ctx)
  "Synchronous: returns (values n xy) with XY a bytevector of N float pairs."
  (let* ((cap (max 1 (%slots ctx)))
         (xy (make-bytevector (* 8 cap)))
         (n (make-i64 1)))
    (check ctx (%positions ctx cap (ptr n) %null-pointer (ptr xy) %null-pointer))
    (values (bytevector-s64-native-ref n 0) xy)))

(define (make-frame-buffer capacity)
  "Page-locked host memory for CAPACITY float pairs, as a bytevector (asynchronous copies need it)."
  (pointer->bytevector (%host-alloc (* 8 capacity)) (* 8 capacity)))

(define (cuda-frame-begin! ctx buffer)
  "Start reading the current frame back into BUFFER (from make-frame-buffer); returns at once.  The
caller may run cuda-step! meanwhile; at most two frames are in flight."
  (check ctx (%frame-begin ctx (ptr buffer) (quotient (bytevector-length buffer) 8))))

(define (cuda-frame-end! ctx)
  "Wait for the oldest frame begun; returns the number of float pairs it holds."
  (let ((n (make-i64 1)))
    (check ctx (%frame-end ctx (ptr n)))
    (bytevector-s64-native-ref n 0)))

;;; Delta frames (include/griddy_cuda.h, "Delta frames"): a frame holds only the actors whose drawn position
;;; changed since the frame before; the painter keeps one bytevector of float pairs per slot and patches it.
(define (make-delta-buffer capacity)
  "Page-locked host memory for one delta frame over CAPACITY slots, as a bytevector."
  (let ((bytes (%delta-bytes capacity)))
    (pointer->bytevector (%host-alloc bytes) bytes)))

(define (cuda-delta-begin! ctx buffer capacity)
  "Start formatting and reading back the delta frame of the current tick into BUFFER; returns at once."
  (check ctx (%delta-begin ctx (ptr buffer) capacity)))

(define (cuda-delta-end! ctx)
  "Wait for the oldest delta frame begun; returns (values slots-covered positions-changed)."
  (let ((n (make-i64 1)) (changed (make-i64 1)))
    (check ctx (%delta-end ctx (ptr n) (ptr changed)))
    (values (bytevector-s64-native-ref n 0) (bytevector-s64-native-ref changed 0))))

(define (cuda-delta-apply! buffer xy)
  "Patch XY, the painter's bytevector of float pairs (one per slot, NaN: nothing to draw), with the frame in BUFFER."
  (let ((rc (%delta-apply (ptr buffer) (ptr xy) (quotient (bytevector-length xy) 8))))
    (unless (zero? rc) (throw 'griddy-cuda rc "griddy_delta_apply: malformed frame or array too small"))))
